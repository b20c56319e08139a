"""oracle/build_ref.py -- TEST INFRASTRUCTURE.  Compiles the reference's OWN Chamfer kernels, from the source where it
lies (/root/reference/extensions/chamfer_distance/chamfer3D.cu, nothing is copied), into oracle/_ref/ref_chamfer3D.cubin
for sm_100a.  Only possible in the build container (the GPU box has no /root/reference); the built file is git-ignored
but travels to the GPU box, where oracle/ref_chamfer.py launches NmDistanceKernel / NmDistanceGradKernel with the
reference's own launch shapes (chamfer3D.cu:142-143, :184-185) as the bit-exactness oracle for dist / idx.

The .cu file's host launchers use at::Tensor, so the torch headers are on the include path; only the device code
(-cubin) is kept."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/extensions/chamfer_distance/chamfer3D.cu"
OUT = os.path.join(HERE, "_ref", "ref_chamfer3D.cubin")


def main():
    if not os.path.exists(SRC):
        print("reference sources not present: keeping the prebuilt oracle/_ref (if any)")
        return 0
    import torch
    from torch.utils.cpp_extension import include_paths
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    inc = []
    for p in include_paths("cuda") if "device_type" in include_paths.__code__.co_varnames else include_paths():
        inc += ["-I", p]
    import sysconfig
    inc += ["-I", sysconfig.get_paths()["include"]]
    cmd = ["/usr/local/cuda/bin/nvcc", "-cubin", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-w",
           "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__"] + inc + [SRC, "-o", OUT]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-2000:])
        return 1
    print("built", OUT)
    return 0


if __name__ == "__main__":
    sys.exit(main())
