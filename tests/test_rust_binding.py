"""rust/azb-sys is the FFI crate north_star asks for.  There is no Rust toolchain in the image, so the crate cannot
be compiled here; what can be checked without one is that it says exactly what include/azb.h says:
  * #[repr(C)] struct fields: same names, order, types and (therefore) offsets as the C structs — compared with the
    ctypes mirror (azdopt_b200/capi.py), whose layout the C compiler's is checked against by azb_create(struct_size);
  * every function the header declares is bound, with the same number of arguments, and nothing else is;
  * every constant has the header's value;
  * azb-nabla calls only functions azb-sys binds.
CPU only."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYS = os.path.join(ROOT, "rust", "azb-sys", "src", "lib.rs")
NABLA = os.path.join(ROOT, "rust", "azb-nabla", "src", "lib.rs")
HEADER = os.path.join(ROOT, "include", "azb.h")

RUST_TYPES = {"u8": (1, 1), "u32": (4, 4), "i32": (4, 4), "u64": (8, 8), "f32": (4, 4), "f64": (8, 8)}


def _strip_comments(src):
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    return re.sub(r"//[^\n]*", " ", src)


def rust_struct(name):
    src = _strip_comments(open(SYS).read())
    m = re.search(r"#\[repr\(C\)\][^{]*?pub struct %s\s*\{(.*?)\n\}" % name, src, flags=re.S)
    assert m, f"struct {name} not found (or not #[repr(C)])"
    fields = []
    for fname, ty in re.findall(r"pub\s+(\w+)\s*:\s*([^,\n]+),", m.group(1)):
        ty = ty.strip()
        arr = re.match(r"\[(\w+);\s*(\d+)\]", ty)
        base, count = (arr.group(1), int(arr.group(2))) if arr else (ty, 1)
        fields.append((fname, base, count))
    return fields


def c_layout(fields):
    """offsets under the C ABI rules #[repr(C)] follows"""
    off, out, align_max = 0, [], 1
    for name, base, count in fields:
        size, align = RUST_TYPES[base]
        off = (off + align - 1) // align * align
        out.append((name, off, size * count))
        off += size * count
        align_max = max(align_max, align)
    return out, (off + align_max - 1) // align_max * align_max


def test_repr_c_structs_match_the_ctypes_mirror():
    from azdopt_b200 import capi

    for rust_name, ct in (("azb_config", capi.Config), ("azb_counters", capi.Counters), ("azb_improvement", capi.Improvement)):
        layout, total = c_layout(rust_struct(rust_name))
        assert [n for n, _, _ in layout] == [n for n, _ in ct._fields_], rust_name
        for name, off, size in layout:
            f = getattr(ct, name)
            assert (f.offset, f.size) == (off, size), (rust_name, name, f.offset, off)
        assert total == C.sizeof(ct), rust_name


def header_functions():
    src = _strip_comments(open(HEADER).read())
    out = {}
    for ret, name, args in re.findall(r"\n\s*((?:const\s+)?\w+\s*\*?)\s*(azb_\w+)\s*\(([^;{]*?)\)\s*;", src):
        args = args.strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def rust_functions():
    src = _strip_comments(open(SYS).read())
    block = re.search(r'extern "C"\s*\{(.*)\}', src, flags=re.S).group(1)
    out = {}
    for name, args in re.findall(r"pub fn (azb_\w+)\s*\(([^)]*)\)", block):
        args = args.strip()
        out[name] = 0 if not args else args.count(":")
    return out


def test_every_header_function_is_bound_with_the_same_arity():
    h, r = header_functions(), rust_functions()
    assert len(h) >= 45
    assert set(h) == set(r), (sorted(set(h) - set(r)), sorted(set(r) - set(h)))
    assert h == r


def test_functions_match_the_exported_symbols_and_ctypes_signatures():
    from azdopt_b200 import capi

    r = rust_functions()
    assert set(r) == set(capi.SIGNATURES)
    for name, (_, args) in capi.SIGNATURES.items():
        assert len(args) == r[name], name


def test_constants_match_the_header():
    hdr = _strip_comments(open(HEADER).read())
    rust = _strip_comments(open(SYS).read())
    consts = dict(re.findall(r"pub const (AZB_\w+)\s*:\s*\w+\s*=\s*([0-9A-Fa-fx_]+)\s*;", rust))
    assert len(consts) >= 17
    for name, val in consts.items():
        v = int(val.replace("_", ""), 0)
        m = re.search(r"#define\s+%s\s+(0x[0-9A-Fa-f]+|\d+)" % name, hdr) or re.search(r"\b%s\s*=\s*(\d+)" % name, hdr)
        assert m, name
        assert int(m.group(1), 0) == v, name


def test_the_optimizer_crate_only_calls_bound_functions():
    used = set(re.findall(r"sys::(azb_\w+)\s*\(", open(NABLA).read()))
    bound = set(rust_functions())
    assert used and used <= bound, sorted(used - bound)
    # the reference's public surface (optimizer/mod.rs:29-363) is all there
    src = open(NABLA).read()
    for method in ("par_new", "par_roll_out_episodes", "argmin_data", "par_update_model", "par_reset_trees", "get_model_mut"):
        assert re.search(r"pub fn %s\b" % method, src), method
    assert "{ … }" not in src and "todo!" not in src and "unimplemented!" not in src
