"""Drop-in ``MAgNetCNN_2d`` (models/magnet_cnn_2d.py:142-478, FACTORY key 'magnet_cnn_2d'): MAgNet with a CNN encoder on a
regular low-resolution grid.  SURVEY §8 f4: its graph stage is textually the same ``Encoder / InteractionNetwork / Processor /
Decoder / _build_graph`` as MAgNet[GNN]'s (models/magnet_cnn_2d.py:13-140,303-327), so it runs on the same kernels here —
radius graph + cached plan (csrc/graph.cu), the fused InteractionNetwork edge kernels (csrc/in_edge_tc.cu,
csrc/in_edge_bwd_tc.cu), the tensor-core MLPs (csrc/linear_tc.cu, csrc/mlp_chain_tc.cu) — by instantiating the classes of
``magnet_b200.magnet_gnn``.  What is specific to the CNN flavour stays ordinary PyTorch and is OUT of the hot path's scope:
the EDSR convolutions (cuDNN) and the nearest-cell ``grid_sample`` of the LIIF-style decoder; the decoder's per-(query, t, corner)
MLP (``proj_head`` + LayerNorm) and the ``projector`` do run on the tensor-core Linears.

Same constructor argument (``hparams``), ``forward`` / ``training_step`` / ``validation_step`` signatures and ``state_dict`` keys
as the reference.  Like MAgNet[GNN] the kernels are built for ``latent_dim = mlp_hidden = n_chan = 128``."""
import torch
from torch import nn
import torch.nn.functional as F

from . import functional as MF
from ._compat import LightningModule
from .magnet_gnn import MLP, Encoder, Processor, Decoder, MAgNetGNN, _mlp_ln


def cell_centres(shape, flatten: bool = True) -> torch.Tensor:
    """Coordinates of the cell centres of a regular grid over [-1, 1]^d, 'ij' order (utils.py:19-36 ``make_coord``)."""
    axes = [(-1.0 + 1.0 / n) + (2.0 / n) * torch.arange(n).float() for n in shape]
    grid = torch.stack(torch.meshgrid(*axes, indexing="ij"), dim=-1)
    return grid.reshape(-1, len(shape)) if flatten else grid


class _ResBlock(nn.Module):
    """conv - ReLU - conv + skip (models/backbones/edsr.py:3-31; the reference's call passes ``res_scale`` into the unused
    ``bias`` slot, so the block's own scale is always 1)."""

    def __init__(self, n_chan, kernel_size):
        super().__init__()
        self.conv_1 = nn.Conv2d(n_chan, n_chan, kernel_size, padding=kernel_size // 2)
        self.conv_2 = nn.Conv2d(n_chan, n_chan, kernel_size, padding=kernel_size // 2)

    def forward(self, x):
        return self.conv_2(F.relu(self.conv_1(x))) + x


class EDSR(nn.Module):
    """EDSR trunk without up-sampling (models/backbones/edsr.py:33-59), 2-D mode; state_dict keys ``head_conv``,
    ``res_layers.{i}.conv_{1,2}``, ``tail_conv``."""

    def __init__(self, in_chan, n_chan=64, res_layers=16, kernel_size=3):
        super().__init__()
        self.head_conv = nn.Conv2d(in_chan, n_chan, kernel_size, padding=kernel_size // 2)
        self.res_layers = nn.Sequential(*[_ResBlock(n_chan, kernel_size) for _ in range(res_layers)])
        self.tail_conv = nn.Conv2d(n_chan, n_chan, kernel_size, padding=kernel_size // 2)
        self.out_dim = n_chan

    def forward(self, x):
        x = self.head_conv(x)
        return self.tail_conv(self.res_layers(x)) + x


class MAgNetCNN_2d(LightningModule):
    """models/magnet_cnn_2d.py:142-478."""

    def __init__(self, hparams):
        super().__init__()
        self.save_hyperparameters()
        for k in ("lr", "weight_decay", "factor", "step_size", "loss", "time_slice", "num_message_passing_steps", "latent_dim",
                  "mlp_layers", "mlp_hidden", "scales", "res_layers", "n_chan", "kernel_size", "res_scale", "interpolation", "radius",
                  "teacher_forcing"):
            setattr(self, k, getattr(hparams, k))
        if not (self.latent_dim == self.mlp_hidden == self.n_chan == 128):
            raise RuntimeError("magnet_b200 kernels are built for latent_dim = mlp_hidden = n_chan = 128; got "
                               f"{self.latent_dim}, {self.mlp_hidden}, {self.n_chan}")
        self.dim = 2
        self.criterion = {"l1": nn.L1Loss(), "l2": nn.MSELoss(), "smooth_l1": nn.SmoothL1Loss()}[self.loss]
        self.mse_criterion, self.mae_criterion = nn.MSELoss(), nn.L1Loss()
        ts, ld = self.time_slice, self.latent_dim
        self.encoder = EDSR(in_chan=ts, n_chan=self.n_chan, res_layers=self.res_layers, kernel_size=self.kernel_size)
        self.proj_head = nn.Sequential(MLP(self.encoder.out_dim + 5 + 1, [self.mlp_hidden] * self.mlp_layers, self.n_chan),
                                       nn.LayerNorm(self.n_chan))
        self.projector = MLP(self.n_chan, [self.mlp_hidden] * self.mlp_layers, 1)
        self._encoder = Encoder(node_in=ts + 3, node_out=ld, edge_in=ts + 2, edge_out=ld, mlp_layers=self.mlp_layers,
                                mlp_hidden=self.mlp_hidden)
        self._processor = Processor(node_in=ld, node_out=ld, edge_in=ld, edge_out=ld,
                                    num_message_passing_steps=self.num_message_passing_steps, mlp_num_layers=self.mlp_layers,
                                    mlp_hidden_dim=self.mlp_hidden)
        self._decoder = Decoder(node_in=ld, node_out=ts, mlp_layers=self.mlp_layers, mlp_hidden=self.mlp_hidden)
        self._graph_cache = {}
        self._grid_cache = {}

    # the graph stage is MAgNet[GNN]'s: same radius graph (rows swapped), node / edge features, cached plan
    _edges = MAgNetGNN._edges
    _build_graph = MAgNetGNN._build_graph

    def _grid(self, W, device):
        hit = self._grid_cache.get((W, str(device)))
        if hit is None:
            hit = self._grid_cache[(W, str(device))] = cell_centres([W, W]).to(device)          # [W*W, 2]
        return hit

    def feature_encoding(self, x_t):
        B, T, C, W, _ = x_t.shape
        return self.encoder(x_t.reshape(B, T * C, W, W))

    def continuous_decoder(self, x_t, feat, cell, coord_hr, t):
        """models/magnet_cnn_2d.py:223-292: for the four neighbouring cells of every query, nearest-cell latent + input values +
        relative position + cell size + time through ``proj_head``, blended by the opposite corner's area.  The reference's T
        separate ``grid_sample`` calls per corner are one call over the T stacked frames; everything else is value-identical."""
        B, C, W, _ = feat.shape
        T, N = x_t.shape[1], coord_hr.shape[1]
        centres = self._grid(W, feat.device).reshape(W, W, 2).permute(2, 0, 1)[None].expand(B, 2, W, W)
        frames = x_t.reshape(B, -1, W, W)                                                # [B, T*Cin, W, W]
        cin = frames.shape[1] // T
        tt = t[:, None, :T].expand(B, N, T)                                              # [B, N, T]
        cell_w = cell * W
        preds, areas = [], []
        for vx in (-1, 1):
            for vy in (-1, 1):
                c = coord_hr.clone()
                c[..., 0] += vx / W + 1e-6
                c[..., 1] += vy / W + 1e-6
                c.clamp_(-1 + 1e-6, 1 - 1e-6)
                grid = c.flip(-1)[:, None]                                               # [B, 1, N, 2] (x, y) order
                samp = lambda src: F.grid_sample(src, grid, mode="nearest", padding_mode="border", align_corners=False)[:, :, 0].permute(0, 2, 1)
                q_feat, q_coord, q_inp = samp(feat), samp(centres), samp(frames)         # [B,N,C], [B,N,2], [B,N,T*Cin]
                rel = (coord_hr - q_coord) * W
                areas.append((rel[..., 0] * rel[..., 1]).abs().reshape(B * N, 1).expand(B * N, T) + 1e-9)
                rows = torch.cat([q_feat[:, :, None].expand(B, N, T, C), q_inp.reshape(B, N, T, cin), rel[:, :, None].expand(B, N, T, 2),
                                  cell_w[:, :, None].expand(B, N, T, 2), tt[..., None]], dim=-1)
                preds.append(_mlp_ln(self.proj_head, rows.reshape(B * N, T, -1)))
        total = areas[0] + areas[1] + areas[2] + areas[3]
        out = 0
        for pred, area in zip(preds, (areas[3], areas[2], areas[1], areas[0])):          # opposite corner's area
            out = out + pred * (area / total)[..., None]
        return out

    def forward(self, x_t, coords, cell, t, hr_last, hiddens=None):
        B, T, _, W, _ = x_t.shape
        N = coords.shape[1]
        T_out = t.shape[-1] - T
        feat = self.feature_encoding(x_t)
        hr_points = self.projector(self.continuous_decoder(x_t, feat, cell, coords, t)).reshape(B, N, -1)     # [B, N, T]
        lr_points = x_t.permute(0, 3, 4, 1, 2).reshape(B, W * W, -1)
        lr_coords = self._grid(W, feat.device)[None].expand(B, W * W, 2)
        all_coords = self._all_coords(lr_coords, coords, W)
        nf, ei, ef, plan = self._build_graph(torch.cat([lr_points, hr_points], 1), all_coords, t[:, :T], return_plan=True)
        nf, ef = self._encoder(nf, ei, ef)
        nf, _ = self._processor(nf, ei, ef, plan=plan, need_e=False)
        ret = self._decoder(nf).reshape(B, W * W + N, -1)
        last = torch.cat([x_t[:, -1].permute(0, 2, 3, 1).reshape(B, W * W, -1), hr_last], dim=1)              # [B, WW+N, 1]
        delta_t = (t[:, T:T + T_out] - t[:, T - 1:T])[:, :, None, None]
        outputs = last[:, None] + delta_t * ret.permute(0, 2, 1)[..., None]                                   # [B, T_out, WW+N, 1]
        out_lr = outputs[:, :, :W * W].permute(0, 1, 3, 2).reshape(B, T_out, -1, W, W)
        return outputs[:, :, W * W:], out_lr, hr_points.reshape(B, N, T, -1).permute(0, 2, 1, 3)

    def _all_coords(self, lr_coords, coords, W):
        """cat([grid, queries]) cached on the query tensor so that the radius graph is built once per mesh."""
        from . import _lib
        key = (coords.data_ptr(), _lib.ver(coords), tuple(coords.shape), W)
        hit = getattr(self, "_coords_cache", None)
        if hit is not None and hit[0] == key and hit[1]() is coords:
            return hit[2]
        import weakref
        cat = torch.cat([lr_coords, coords], dim=1)
        self._coords_cache = (key, weakref.ref(coords), cat)
        return cat

    configure_optimizers = MAgNetGNN.configure_optimizers

    def _unroll(self, batch, teacher_forcing: bool, validation: bool):
        t, u, uv = batch["t"].float(), batch["lr_frames"].float(), batch["hr_points"].float()
        coords, cells = batch["coords"].float(), batch["cells"].float()
        B, _, _, W, _ = u.shape
        ts = self.time_slice
        T_future = uv.shape[1] - ts
        inp, hr_last = u[:, :ts], uv[:, ts - 1]
        preds, hr_preds = [], []
        for i in range(T_future // ts):
            out_hr, out_lr, hr_points = self.forward(inp, coords, cells, t[:, i * ts:(i + 2) * ts], hr_last)
            hr_preds.append(hr_points)
            if validation:
                # models/magnet_cnn_2d.py:452-466: only the query predictions are kept, and they are fed back as the next
                # low-resolution frames through a bilinear resize (the query set is a square grid there)
                preds.append(out_hr)
                g = out_hr.permute(0, 1, 3, 2)
                w_in = int(g.shape[-1] ** 0.5)
                g = g.reshape(-1, g.shape[2], w_in, w_in)
                inp = F.interpolate(g, size=W, mode="bilinear", align_corners=False).reshape(B, out_hr.shape[1], -1, W, W)
                hr_last = out_hr[:, -1]
            else:
                preds.append(torch.cat([out_hr, out_lr.reshape(*out_lr.shape[:3], -1).permute(0, 1, 3, 2)], dim=2))
                if teacher_forcing:
                    inp, hr_last = u[:, (i + 1) * ts:(i + 2) * ts], uv[:, (i + 2) * ts - 1]
                else:
                    inp, hr_last = out_lr, out_hr[:, -1]
        return torch.cat(preds, dim=1), torch.cat(hr_preds, dim=1)

    def training_step(self, train_batch, batch_idx):
        u, uv = train_batch["lr_frames"].float(), train_batch["hr_points"].float()
        B, T, C = u.shape[:3]
        ts = self.time_slice
        pred, hr_pred = self._unroll(train_batch, self.teacher_forcing, validation=False)
        target = torch.cat([uv[:, ts:], u[:, ts:].reshape(B, T - ts, C, -1).permute(0, 1, 3, 2)], dim=2)
        loss = self.criterion(pred, target) + self.criterion(hr_pred, uv[:, :-ts])
        self.log("train_loss", loss, prog_bar=True)
        self.log("train_mae_loss", self.mae_criterion(pred, target), prog_bar=True)
        self.log("train_interp_loss", self.mae_criterion(hr_pred, uv[:, :-ts]), prog_bar=True)
        return loss

    def validation_step(self, val_batch, batch_idx):
        uv = val_batch["hr_points"].float()
        pred, _ = self._unroll(val_batch, False, validation=True)
        self.log("val_loss", self.criterion(pred, uv[:, self.time_slice:]), prog_bar=True)
        self.log("val_mae_loss", self.mae_criterion(pred, uv[:, self.time_slice:]), prog_bar=True)
