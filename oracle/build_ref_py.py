"""oracle/build_ref_py.py -- TEST / BASELINE INFRASTRUCTURE.  Byte-compiles the reference's OWN Python implementation of the hot path,
from the sources where they lie under /root/reference (nothing is copied into the repository), into sourceless byte code packed in the
git-ignored archive oracle/_ref/refpy.zip -- the Python counterpart of oracle/build_ref.py's cubin (an archive because the GPU-box
snapshot drops loose *.pyc files).  Only possible in the build container (the GPU
box has no /root/reference); the built files travel to the GPU box, where oracle/ref_model.py imports them:

  models/**                                   models/model.py:9-64 PCNNet, models/pcn.py, models/vn_layers.py, ... (the whole package:
                                              models/__init__.py imports all of it)
  metrics/loss.py                             cd_loss_L1 / cd_loss_L2 (metrics/loss.py:20-43)
  extensions/ChamferDistancePytorch/chamfer_python.py, fscore.py
                                              distChamfer, the reference's CPU-capable Chamfer (chamfer_python.py:18-39)
  utils/loss.py                               calc_cd / calc_dcd
  extensions/chamfer_distance/chamfer_distance.py
                                              the reference's autograd wrapper (:29-84); on the GPU box it runs on top of the reference's own
                                              kernels (oracle/_ref/ref_chamfer3D.cubin, launched by oracle/ref_chamfer.py)

    python oracle/build_ref_py.py

The .pyc files are tied to this interpreter's magic number (the GPU box runs the same image)."""
import os
import py_compile
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
OUT = os.path.join(HERE, "_ref", "refpy.zip")

TREES = ["models"]
FILES = ["metrics/loss.py", "utils/loss.py", "extensions/ChamferDistancePytorch/chamfer_python.py",
         "extensions/ChamferDistancePytorch/fscore.py", "extensions/chamfer_distance/chamfer_distance.py"]


def main():
    if not os.path.isdir(REF):
        print("reference sources not present: keeping the prebuilt oracle/_ref/refpy.zip (if any)")
        return 0
    todo = list(FILES)
    for tree in TREES:
        for root, _, files in os.walk(os.path.join(REF, tree)):
            for f in files:
                if f.endswith(".py"):
                    todo.append(os.path.relpath(os.path.join(root, f), REF))
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    tmp = tempfile.mkdtemp(prefix="refpy_")
    try:
        dirs = set()
        with zipfile.ZipFile(OUT, "w", zipfile.ZIP_STORED) as z:
            for rel in sorted(todo):
                src = os.path.join(REF, rel)
                dst = os.path.join(tmp, rel[:-3] + ".pyc")
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                # dfile = the reference path, so tracebacks cite the reference's own file:line
                py_compile.compile(src, cfile=dst, dfile=src, doraise=True, optimize=0,
                                   invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
                d = os.path.dirname(rel)
                while d and d not in dirs:      # explicit directory entries: zipimport needs them for packages without __init__
                    dirs.add(d)
                    z.writestr(d + "/", b"")
                    d = os.path.dirname(d)
                z.write(dst, rel[:-3] + ".pyc")
            z.writestr("MANIFEST.txt", f"byte-compiled from {REF} by oracle/build_ref_py.py with python {sys.version.split()[0]}\n"
                       + "\n".join(sorted(todo)) + "\n")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    old = os.path.join(HERE, "_ref", "py")
    if os.path.isdir(old):
        shutil.rmtree(old)
    print(f"compiled {len(todo)} reference modules into {OUT}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
