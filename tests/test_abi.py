"""CPU-only checks of the drop-in boundary: libdlra.so loads without a GPU and exports every symbol that include/dlra.h
declares (no compute calls here), and the ctypes table covers the whole header."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "dlra.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dlra_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    import lowrankintegrators.jl_b200 as lri
    lib = ctypes.CDLL(lri._lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"libdlra.so does not export {s}"
    assert set(syms) == set(lri._lib.SIGNATURES), set(syms) ^ set(lri._lib.SIGNATURES)
    assert lri._lib.load().dlra_version().startswith(b"dlra-b200")


def test_no_cpu_fallback_create_fails_loudly_without_gpu():
    import torch
    import lowrankintegrators.jl_b200 as lri
    if torch.cuda.is_available():
        return
    try:
        lri.Engine(64, 32, 4)
    except lri._lib.DLRAError as e:
        assert e.code != 0
    else:
        raise AssertionError("Engine creation must fail without a CUDA device")


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lowrankintegrators.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_header_is_plain_c99_and_links_from_c(tmp_path):
    """The boundary is a C ABI: include/dlra.h must compile as strict C99 (no C++/torch types) and a C program must link
    against libdlra.so and reach an entry point that needs no GPU."""
    import shutil
    import subprocess
    import lowrankintegrators.jl_b200 as lri
    if shutil.which("gcc") is None:
        return
    src = tmp_path / "abi_probe.c"
    src.write_text('#include <stdio.h>\n#include <string.h>\n#include "dlra.h"\n'
                   'int main(void) {\n'
                   '  dlra_handle h = 0;\n'
                   '  if (strncmp(dlra_version(), "dlra-b200", 9) != 0) return 2;\n'
                   '  if (dlra_create(0, 0, 0, 0, 0, 0, &h) == 0) return 3;   /* invalid sizes are refused before any CUDA call */\n'
                   '  if (h != 0 || strlen(dlra_last_error(0)) == 0) return 4;\n'
                   '  puts("abi ok");\n  return 0;\n}\n')
    exe = tmp_path / "abi_probe"
    libdir = os.path.dirname(lri._lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-l:libdlra.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "abi ok" in out.stdout, (out.returncode, out.stdout, out.stderr)
