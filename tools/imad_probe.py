import sys; sys.path.insert(0,'.')
from circom_cvm_b200 import engine as E
names=["lo+hi pair","mad.wide","2x mad.lo","2x mad.hi","2x addc chain","2x add","2x madc chain","mul.wide+add","IMAD.WIDE.X chain"]
for k,n in enumerate(names):
    v,ms=E.imad_peak(k); print("%-14s %.2f Tunits/s  (%.1f units/clk/SM @1.965GHz) %.3f ms"%(n,v/1e12,v/148/1.965e9,ms))
for variant,name in ((0,"portable"),(1,"mul.wide 2 chains"),(2,"mul.wide 1 chain"),(3,"carry-chained"),(4,"squaring")):
    for c in (1,2,4,5,6,8,12,16):
        v=E.mul_peak(variant,c); print("mont_mul %-17s %d CTA/SM (%2d warps): %.2f Gmul/s = %.2f clk*SM per mul, %.2f Tmac/s"%(name,c,4*c,v/1e9,148*1.965e9/v,v*136/1e12))
