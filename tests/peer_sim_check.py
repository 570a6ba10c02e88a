"""Child process of tests/test_gpu_peer_reduce.py::test_simulated_ranks_match_full_batch_update: the simulated-rank cases of the
peer-memory Adam kernel on one GPU.  Prints 'case ok' per case and 'peer-sim ok' at the end."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from test_gpu_peer_reduce import SIM_CASES, simulated_ranks_case

if __name__ == '__main__':
    for system, world, static in SIM_CASES:
        simulated_ranks_case(system, world, static)
        print('case ok', system, world, static, flush=True)
    print('peer-sim ok', flush=True)
