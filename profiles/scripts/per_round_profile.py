"""cProfile of the PER round (sample + update_priorities) through the Python mirror at B = 4096 / 64: where the host time goes."""
import sys, random, cProfile, pstats, io; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from types import SimpleNamespace
from cacto_b200.replay_buffer import PrioritizedReplayBuffer
ns = 7
rng = np.random.default_rng(0)
for B in (4096, 64):
    conf = SimpleNamespace(REPLAY_SIZE=2**16, BATCH_SIZE=B, nb_state=ns, prioritized_replay_alpha=0.6, prioritized_replay_beta=0.6,
                           prioritized_replay_eps=1e-2, fresh_factor=0.95)
    rows = rng.normal(size=(2**16 + 8, 3 * ns + 3))
    cols = (rows[:, :ns], rows[:, ns], rows[:, ns + 1:2 * ns + 1], rows[:, 2 * ns + 1:3 * ns + 1], rows[:, 3 * ns + 1], rows[:, 3 * ns + 2])
    pb = PrioritizedReplayBuffer(conf); pb.add(*[(c,) for c in cols])
    rtg = torch.randn(B, 1, device='cuda'); V = torch.randn(B, 1, device='cuda')
    random.seed(0)
    for _ in range(10):
        o = pb.sample(); pb.update_priorities(o[7], rtg, V)
    torch.cuda.synchronize()
    pr = cProfile.Profile(); pr.enable()
    for _ in range(200):
        o = pb.sample(); pb.update_priorities(o[7], rtg, V)
    torch.cuda.synchronize(); pr.disable()
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22)
    print(f'==== B = {B} (200 rounds; divide by 200)'); print('\n'.join(l[:150] for l in s.getvalue().splitlines()[4:40]))
