"""TEST INFRASTRUCTURE: builds the reference's own second-derivative sampler kernel into oracle/_ref/.

The only native code of the reference on the hot path is its torch extension
``custom/triplaneturbo/extern/grid_sample_gradfix/gridsample_cuda.{cpp,cu}`` (pybind module ``gridsample_grad2``:
``grad2_2d``, ``grad2_3d``).  It is compiled HERE from the sources where they lie under /root/reference (nothing is
copied into the repository) with torch's own extension builder (nvcc + ninja, sm_100a), and only the resulting
``oracle/_ref/gridsample_grad2_ref.so`` travels to the GPU box (git-ignored, not gpurun-ignored).  There it serves one
purpose: ``tests/test_sampler.py::test_second_derivative_against_the_reference_kernel`` checks
``tt_sample_planes_bwdbwd`` against the reference's kernel on identical inputs.  The product never loads it.

    python oracle/build_ref.py         (called by __graft_entry__.build() when /root/reference is present)
"""
import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/custom/triplaneturbo/extern/grid_sample_gradfix"
OUT = os.path.join(HERE, "_ref")
NAME = "gridsample_grad2_ref"


def build(verbose=False):
    if not os.path.isdir(REF):
        return None
    target = os.path.join(OUT, NAME + ".so")
    srcs = [os.path.join(REF, "gridsample_cuda.cpp"), os.path.join(REF, "gridsample_cuda.cu")]
    if os.path.exists(target) and all(os.path.getmtime(target) >= os.path.getmtime(s) for s in srcs):
        return target
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    build_dir = os.path.join(OUT, "build")
    os.makedirs(build_dir, exist_ok=True)
    load(name=NAME, sources=srcs, build_directory=build_dir, verbose=verbose, is_python_module=False,
         extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a"])
    so = glob.glob(os.path.join(build_dir, NAME + "*.so"))
    shutil.copy(so[0], target)
    shutil.rmtree(build_dir, ignore_errors=True)
    return target


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
