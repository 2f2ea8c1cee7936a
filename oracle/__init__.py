"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the hot path of jaggbow/magnet (the reference) and of the
un-vendored third-party functions it calls.  The product (``magnet_b200/``)
never imports this package: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs do, and only as the
checker / reported baseline.

PARITY UNPINNED: the reference ships no tests, golden vectors
or known-answer files, and its neighbour-search / scatter / norm arithmetic lives
in torch_cluster / torch_scatter / torch_geometric, none of which is installable
here.  What pins this oracle instead:

* ``oracle/thirdparty/`` restates those third-party functions (semantics recorded in
  SURVEY.md §8c) and is self-checked against scipy's cKDTree, dense one-hot
  matmuls and ``torch.nn.functional.instance_norm`` (tests/test_oracle_*.py);
* ``oracle/reference_loader.py`` imports the reference's UNMODIFIED model files
  (from /root/reference, this container only) on top of those stubs;
  ``oracle/gen_golden.py`` runs them on seeded inputs and commits the outputs under
  ``tests/golden/``;
* ``oracle/restatement.py`` — the part that travels to the GPU box — restates the
  reference's own model code for the path and is checked against those golden
  vectors.
"""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libgraph_oracle.so")


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc). Building the checker is not using it."""
    src = os.path.join(_HERE, "csrc", "graph_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "all"])
    return LIB_PATH
