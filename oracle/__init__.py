"""CPU oracle for the ST-GCN / RT-ST-GCN forward path.

TEST INFRASTRUCTURE ONLY.  This package is a CPU restatement of the reference
algorithm (maximyudayev/Realtime-ST-GCN) used as the *checker* for the CUDA
path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``realtime-st-gcn_b200/`` imports it, and the product path has no CPU fallback.

Parity pinning: the reference ships no golden vectors, so the
oracle is pinned against outputs of the reference modules themselves, generated
in the build container by ``tools/make_golden.py`` (which imports
``/root/reference``) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every fixture.
"""
from .graph_oracle import build_adjacency  # noqa: F401
from . import stgcn_oracle  # noqa: F401
