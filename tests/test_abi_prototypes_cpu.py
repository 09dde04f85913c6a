"""The ctypes binding (monica_b200/_lib.py) against the header it binds (include/monica_b200.h), mechanically: every
prototype's argument count and argument / return kinds (pointer, 32-bit int, 64-bit int, float, double), and the four
structures' field offsets and sizes as the C compiler lays them out.  A mismatch here is memory corruption on the GPU box,
not a failed assertion, so it is checked where gcc is."""
import ctypes as C
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "monica_b200.h")


def _c_kind(t):
    t = re.sub(r"\bconst\b", "", t).strip()
    if "*" in t or "[" in t:
        return "ptr"
    t = t.split()[0] if t.split() else t
    return {"int": "i32", "int32_t": "i32", "uint32_t": "i32", "int8_t": "i8", "uint8_t": "i8", "int64_t": "i64", "uint64_t": "i64",
            "float": "f32", "double": "f64", "void": "void"}[t]


def _ct_kind(t):
    if t is None:
        return "void"
    if t in (C.c_void_p, C.c_char_p) or hasattr(t, "contents") or issubclass(t, C._Pointer):
        return "ptr"
    size = C.sizeof(t)
    if t in (C.c_float,):
        return "f32"
    if t in (C.c_double,):
        return "f64"
    return {1: "i8", 4: "i32", 8: "i64"}[size]


def _prototypes():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \t]*?[\s\*]+)\b(mb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        kinds = []
        for a in params:
            if "*" in a or "[" in a:
                kinds.append("ptr")
            else:
                kinds.append(_c_kind(" ".join(a.split()[:-1])))      # drop the parameter name
        out[name] = (_c_kind(ret), kinds)
    return out


def test_every_ctypes_prototype_matches_the_header():
    from monica_b200 import _lib
    L = _lib.lib()
    protos = _prototypes()
    assert set(protos) == set(_lib.SYMBOLS)
    bad = []
    for name, (ret, kinds) in sorted(protos.items()):
        fn = getattr(L, name)
        got_ret = _ct_kind(fn.restype)
        if got_ret != ret:
            bad.append(f"{name}: returns {ret} in the header, {got_ret} in ctypes")
        if fn.argtypes is None:
            if kinds:
                bad.append(f"{name}: {len(kinds)} parameters in the header, no argtypes in ctypes")
            continue
        got = [_ct_kind(t) for t in fn.argtypes]
        if got != kinds:
            bad.append(f"{name}: header {kinds} vs ctypes {got}")
    assert not bad, "\n".join(bad)


def test_structures_have_the_compiler_s_layout(tmp_path):
    from monica_b200 import _lib
    structs = {"mb_opt_t": _lib.Opt, "mb_stats_t": _lib.Stats, "mb_dp_task_t": _lib.DpTask, "mb_ll_task_t": _lib.LLTask}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["return 0; }"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-o", str(exe), str(src)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    want = {}
    for ln in out.splitlines():
        s, f, v = ln.split()
        want[(s, f)] = int(v)
    for cname, ct in structs.items():
        assert C.sizeof(ct) == want[(cname, "size")], cname
        for fname, _ in ct._fields_:
            assert getattr(ct, fname).offset == want[(cname, fname)], (cname, fname)
