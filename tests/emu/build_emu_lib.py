"""Build libdfx_emu_full.so: the WHOLE C-ABI library (include/dfx.h) compiled by g++ from the
product sources (distributed_faiss_b200/csrc/*.cu with -DDFX_EMU) on top of the fiber SIMT runtime
(simt.h) and the CUDA stand-in headers (shim/).  Kernels run one CTA after the other on the CPU;
the tcgen05 screening kernel of the coarse quantizer is replaced by a plain C++ restatement of its
result (dfx_tc.cu, DFX_EMU branch); group selection, exact re-evaluation and the drivers are the product code.  Test infrastructure: nothing in the product loads this file.

    python tests/emu/build_emu_lib.py          # prints the path of the library
    DFX_EMU_LIB=$(python tests/emu/build_emu_lib.py) python -m pytest tests -m gpu -k "<small cases>"

With --asan the library is instrumented by AddressSanitizer: "device memory" is malloc'ed host
memory in this build, so an out-of-bounds global or shared access of any kernel, or of the host
drivers, aborts the test -- the CPU stand-in for `compute-sanitizer --tool memcheck`:
    LD_PRELOAD=$(g++ -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    DFX_EMU_LIB=$(python tests/emu/build_emu_lib.py --asan) python -m pytest tests -m gpu -k "<small cases>"
"""
import os
import shutil
import subprocess
import sys

EMU_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(EMU_DIR))
CSRC = os.path.join(ROOT, "distributed_faiss_b200", "csrc")
OUT_DIR = os.path.join(EMU_DIR, "_build", "lib")
LIB = os.path.join(EMU_DIR, "_build", "libdfx_emu_full.so")
SOURCES = ["dfx_api.cu", "dfx_search.cu", "dfx_build.cu", "dfx_scan_il.cu", "dfx_scan_il2.cu", "dfx_tc.cu"]
FLAGS = ["-std=c++17", "-O1", "-g", "-fPIC", "-ffp-contract=off", "-DDFX_EMU", "-Wno-unknown-pragmas",
         "-Wno-attributes", "-I", os.path.join(EMU_DIR, "shim"), "-I", CSRC]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    d += [os.path.join(EMU_DIR, "simt.h"), os.path.join(ROOT, "include", "dfx.h")]
    for base, _, files in os.walk(os.path.join(EMU_DIR, "shim")):
        d += [os.path.join(base, f) for f in files]
    return d


def build(verbose=False, asan=False):
    cxx = shutil.which("g++")
    if cxx is None:
        raise RuntimeError("g++ not available")
    out_dir = OUT_DIR + ("_asan" if asan else "")
    lib = LIB.replace(".so", "_asan.so") if asan else LIB
    flags = FLAGS + (["-fsanitize=address", "-fno-omit-frame-pointer"] if asan else [])
    os.makedirs(out_dir, exist_ok=True)
    dep_time = max(os.path.getmtime(p) for p in _deps())
    objs, relink = [], not os.path.exists(lib)
    jobs = []
    for src in SOURCES + ["emu_simt_impl.cpp"]:
        path = os.path.join(CSRC, src) if src.endswith(".cu") else os.path.join(EMU_DIR, src)
        obj = os.path.join(out_dir, os.path.splitext(src)[0] + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(dep_time, os.path.getmtime(path)):
            cmd = [cxx] + (["-x", "c++"] if src.endswith(".cu") else []) + flags + ["-c", path, "-o", obj]
            jobs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
            relink = True
    for j in jobs:
        out, _ = j.communicate()
        if j.returncode != 0:
            raise RuntimeError("emulator build failed:\n" + out[-4000:])
        if verbose and out:
            print(out, file=sys.stderr)
    if relink:
        subprocess.run([cxx, "-shared"] + (["-fsanitize=address"] if asan else []) + ["-o", lib] + objs, check=True)
    return lib


if __name__ == "__main__":
    print(build(verbose=True, asan="--asan" in sys.argv[1:]))
