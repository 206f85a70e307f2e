#!/usr/bin/env python
"""Generate the `extern "C"` block of rust/kmers-b200-sys/src/lib.rs from include/kmers_b200.h, so that the Rust view of
the C ABI cannot drift from the header (there is no Rust toolchain in this image to catch it at compile time).

  python scripts/gen_rust_sys.py           rewrite the block between the BEGIN/END GENERATED markers
  python scripts/gen_rust_sys.py --check   exit 1 if the file differs from what would be generated

tests/test_abi.py::test_rust_sys_matches_the_header runs the same comparison."""
from __future__ import annotations

import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "kmers_b200.h")
LIB_RS = os.path.join(ROOT, "rust", "kmers-b200-sys", "src", "lib.rs")
BEGIN, END = "    // BEGIN GENERATED (scripts/gen_rust_sys.py)\n", "    // END GENERATED\n"

SCALARS = {"int32_t": "i32", "uint32_t": "u32", "uint64_t": "u64", "int64_t": "i64", "uint8_t": "u8", "uint16_t": "u16", "size_t": "usize",
           "double": "f64", "char": "c_char", "void": "c_void", "kmb_ctx": "kmb_ctx", "kmb_digest": "kmb_digest"}
KEYWORDS = {"in": "input", "type": "ty", "ref": "reference", "match": "matched"}


def prototypes(header_text: str):
    """[(name, return C type, [(C type, param name)])] for every kmb_* function declared in the header."""
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    text = re.sub(r"#.*", "", text)
    out = []
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(kmb_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = []
        if params and params != "void":
            for p in params.split(","):
                p = " ".join(p.split())
                pm = re.match(r"(.*?)(\w+)$", p)
                plist.append((pm.group(1).strip(), pm.group(2)))
        out.append((name, ret, plist))
    return out


def rust_type(c: str) -> str:
    """C type -> Rust FFI type.  Pointers are read right to left: `kmb_ctx *const *` = const pointer to mutable pointer."""
    toks = c.replace("*", " * ").split()
    # base type with an optional leading const
    base_const = toks[0] == "const"
    if base_const:
        toks = toks[1:]
    base, rest = toks[0], toks[1:]
    ty = SCALARS[base]
    pointee_const = base_const
    for t in rest:
        if t == "*":
            ty = ("*const " if pointee_const else "*mut ") + ty
            pointee_const = False
        elif t == "const":
            pointee_const = True  # applies to the pointer just formed: the NEXT level sees a const pointee
        else:
            raise ValueError(f"cannot map C type {c!r}")
    return ty


def generate(header_text: str) -> str:
    lines = []
    for name, ret, params in prototypes(header_text):
        ps = ", ".join(f"{KEYWORDS.get(n, n)}: {rust_type(t)}" for t, n in params)
        r = "" if ret == "void" else f" -> {rust_type(ret)}"
        lines.append(f"    pub fn {name}({ps}){r};\n")
    return "".join(lines)


def render(lib_rs_text: str, block: str) -> str:
    a, b = lib_rs_text.index(BEGIN) + len(BEGIN), lib_rs_text.index(END)
    return lib_rs_text[:a] + block + lib_rs_text[b:]


def main():
    with open(HEADER) as f:
        block = generate(f.read())
    with open(LIB_RS) as f:
        cur = f.read()
    new = render(cur, block)
    if "--check" in sys.argv:
        sys.exit(0 if new == cur else 1)
    with open(LIB_RS, "w") as f:
        f.write(new)
    print(f"{LIB_RS}: {block.count('pub fn')} entry points")


if __name__ == "__main__":
    main()
