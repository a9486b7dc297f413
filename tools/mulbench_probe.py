import sys; sys.path.insert(0,'.')
from circom_cvm_b200 import engine as E
for variant in (2, 3, 4):
    print(variant, E.mul_peak(variant, 5))
