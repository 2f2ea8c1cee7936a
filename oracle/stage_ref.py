"""ORACLE (test infrastructure only) — stage the reference's OWN implementation of the path for the GPU box.

The reference is pure Python: "building" it means byte-compiling its model files, from the sources where they
lie under /root/reference, into oracle/_ref/ (git-ignored, travels with gpurun like the built .so files).  No reference
SOURCE enters this repository or the snapshot: only CPython bytecode (.pyc format, sourceless imports through oracle/reference_loader.py), the analogue of a compiled
reference binary.  The GPU box runs the same image (same CPython), so the bytecode loads there.

    python -m oracle.stage_ref            # also run by __graft_entry__.build() when /root/reference is present

Staged: models/{mpnn,mpnn_2d,magnet_gnn,magnet_cnn_2d}.py, models/backbones/{mlp,edsr}.py, utils.py — the files SURVEY §8(a)
cites plus the MAgNet[CNN] model whose GNN stage reuses the same classes (§8 f4).  They import on top of oracle/thirdparty/
(restated torch_geometric / pytorch_lightning surface).  bench.py --impl reference and cpu_baseline time THESE modules.
"""
import os
import py_compile
import shutil
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("MAGNET_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(_HERE, "_ref")
EXT = ".pyc.bin"      # not "*.pyc": snapshot tools drop those; reference_loader imports these files by explicit path
FILES = ["utils.py", "models/mpnn.py", "models/mpnn_2d.py", "models/magnet_gnn.py", "models/magnet_cnn_2d.py",
         "models/backbones/mlp.py", "models/backbones/edsr.py"]


def stage(force: bool = False) -> str:
    if not os.path.isfile(os.path.join(REF_ROOT, "models", "mpnn_2d.py")):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    tag = os.path.join(OUT, "STAGED")
    stamp = f"{sys.version_info[:3]} " + " ".join(f"{f}:{os.path.getmtime(os.path.join(REF_ROOT, f)):.0f}" for f in FILES)
    if not force and os.path.exists(tag) and open(tag).read() == stamp:
        return OUT
    shutil.rmtree(OUT, ignore_errors=True)
    for f in FILES:
        dst = os.path.join(OUT, f[:-3] + EXT)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile: the path recorded in tracebacks points at the reference tree, not at this repository
        py_compile.compile(os.path.join(REF_ROOT, f), cfile=dst, dfile=os.path.join("reference", f), doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    open(tag, "w").write(stamp)
    return OUT


if __name__ == "__main__":
    print(stage(force="--force" in sys.argv))
