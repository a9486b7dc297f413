"""include/cvmgpu.h -> the Rust `extern "C"` binding shown in INTEGRATION.md (struct layouts and every prototype).
Run `python tools/gen_rust_ffi.py` to print it; tests/test_abi.py checks that INTEGRATION.md holds exactly this text, so the
binding a maintainer copies cannot drift from the header (round 1's ProgramInfo did)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CTYPES = {"uint64_t": "u64", "uint32_t": "u32", "uint8_t": "u8", "int": "c_int", "size_t": "usize", "double": "f64", "void": "c_void",
          "char": "c_char", "cvmgpu_program": "CvmgpuProgram", "cvmgpu_r1cs": "CvmgpuR1cs", "cvmgpu_program_info": "ProgramInfo",
          "cvmgpu_r1cs_info": "R1csInfo"}
STRUCTS = {"cvmgpu_program_info": "ProgramInfo", "cvmgpu_r1cs_info": "R1csInfo"}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def rust_type(ctype):
    ctype = ctype.strip()
    const = ctype.startswith("const ")
    base = ctype.replace("const ", "").strip()
    stars = base.count("*")
    base = base.replace("*", "").strip()
    t = CTYPES[base]
    for _ in range(stars):
        t = ("*const " if const else "*mut ") + t
        const = False if stars > 1 else const       # `const T **out`: the inner pointer is const, the outer one is written
    if ctype.replace(" ", "").endswith("**") and ctype.startswith("const "):
        t = "*mut *const " + CTYPES[base]
    return t


def generate():
    text = strip_comments(open(os.path.join(ROOT, "include", "cvmgpu.h")).read())
    out = ["use std::os::raw::{c_char, c_int, c_void};", "",
           "#[repr(C)] pub struct CvmgpuProgram { _p: [u8; 0] }", "#[repr(C)] pub struct CvmgpuR1cs { _p: [u8; 0] }", ""]
    for cname, rname in STRUCTS.items():
        end = text.index("} %s;" % cname)
        body = text[text.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        fields = []
        for decl in body.split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            ctype, names = decl.split(" ", 1)
            for n in names.split(","):
                fields.append("    pub %s: %s," % (n.strip(), CTYPES[ctype]))
        out += ["// set struct_size = size_of::<%s>() as u32 before the call; the library fills at most that many bytes" % rname,
                "#[repr(C)] #[derive(Default, Debug, Clone)]", "pub struct %s {" % rname] + fields + ["}", ""]
    out.append('extern "C" {')
    protos = re.findall(r"^([A-Za-z_][\w \*]*?)\b(cvmgpu_\w+)\(([^;]*?)\);", text, flags=re.M | re.S)
    for ret, name, args in protos:
        ret = ret.strip()
        params = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                m = re.match(r"(.*?)(\w+)$", a)
                params.append("%s: %s" % (m.group(2), rust_type(m.group(1))))
        r = "" if ret == "void" else " -> %s" % rust_type(ret)
        out.append("    pub fn %s(%s)%s;" % (name, ", ".join(params), r))
    out.append("}")
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    print(generate(), end="")
