"""TEST INFRASTRUCTURE ONLY -- build the REFERENCE's own CUDA sampler (sampler/sampler.cpp +
sampler/sampler_kernel.cu) for sm_100a into oracle/_ref/ (git-ignored; travels to the GPU box).

The sources are compiled from where they lie under /root/reference; the only change is the two-token patch the
reference needs to compile against torch >= 2.x (`volume.type()` -> `volume.scalar_type()` in the two
AT_DISPATCH lines, sampler_kernel.cu:126,157), applied to a scratch copy under /tmp.  Nothing is copied into the
repository.  The result is the CUDA oracle and the "kernel to beat" for corr_sampler (tests/test_gpu_ref_sampler.py).
"""
import os
import re
import shutil
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("ANYSTEREO_REFERENCE", "/root/reference") + "/sampler"
OUT = os.path.join(ROOT, "oracle", "_ref")


def main():
    if not os.path.isdir(SRC):
        print("reference sampler sources not present; nothing to do")
        return 0
    os.makedirs(OUT, exist_ok=True)
    if any(f.startswith("corr_sampler_ref") and f.endswith(".so") for f in os.listdir(OUT)):
        print("oracle/_ref already built")
        return 0
    tmp = tempfile.mkdtemp(prefix="anystereo_ref_sampler_")
    for f in ("sampler.cpp", "sampler_kernel.cu"):
        shutil.copy(os.path.join(SRC, f), os.path.join(tmp, f))
    p = os.path.join(tmp, "sampler_kernel.cu")
    s = open(p).read()
    s, n = re.subn(r"AT_DISPATCH_FLOATING_TYPES_AND_HALF\(volume\.type\(\)", "AT_DISPATCH_FLOATING_TYPES_AND_HALF(volume.scalar_type()", s)
    assert n == 2, n
    open(p, "w").write(s)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    build = os.path.join(tmp, "build")
    os.makedirs(build, exist_ok=True)
    load(name="corr_sampler_ref", sources=[os.path.join(tmp, "sampler.cpp"), p], build_directory=build,
         extra_cuda_cflags=["-O3", "-lineinfo"], verbose=False, is_python_module=False)
    for f in os.listdir(build):
        if f.endswith(".so"):
            shutil.copy(os.path.join(build, f), os.path.join(OUT, f))
            print("built", os.path.join(OUT, f))
    shutil.rmtree(tmp, ignore_errors=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
