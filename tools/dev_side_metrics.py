"""Dev helper (GPU): the INR-decode and rollout side metrics of bench.py on their own (for ncu launch lists).
usage: python tools/dev_side_metrics.py [decode|rollout] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from magnet_b200 import synthetic as S
from magnet_b200.magnet_gnn import MAgNetGNN

what = sys.argv[1] if len(sys.argv) > 1 else "decode"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)


class HP(dict):
    __getattr__ = dict.__getitem__


hp = HP(time_slice=10, latent_dim=128, num_message_passing_steps=5, mlp_layers=4, mlp_hidden=128, radius=0.08, n_chan=128,
        teacher_forcing=True, codec_neighbors=4, noise=0, interpolation="area", factor=0.3, step_size=50, loss="l1",
        lr=1e-3, weight_decay=0)
m = MAgNetGNN(hp).to(dev).eval()
m.load_state_dict(S.seeded_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 7))
with torch.no_grad():
    if what == "decode":
        g = S._gen(500)
        Lr, Q, T = 1 << 18, 1 << 18, 10
        lr_coords = (2 * torch.rand(1, Lr, 2, generator=g) - 1).to(dev)
        hr_coords = (2 * torch.rand(1, Q, 2, generator=g) - 1).to(dev)
        enc = torch.randn(1, Lr, 128, generator=g).to(dev)
        x_lr = torch.randn(1, T, 1, Lr, generator=g).to(dev)
        t = torch.linspace(0, 1, 2 * T)[None].to(dev)
        fn = lambda: m.projector(m.continuous_decoder(x_lr, enc, lr_coords, hr_coords, t))
    else:
        b = {k: v.to(dev) for k, v in S.implicit_batch(B=32, L=256, Nq=256, nt=50, d=2, kind="concentrated", seed=600).items()}
        fn = lambda: m.rollout(b, teacher_forcing=False)
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    print(what, (time.perf_counter() - t0) / reps * 1e3, "ms per call")
