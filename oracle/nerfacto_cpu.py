"""CPU restatement of the reference's own CPU-runnable path -- BASELINE.json configs[0]: "nerfstudio nerfacto with
torch hash-encoding backend, 4096 synthetic rays x 48 samples, fwd+bwd on CPU (reference path, no GPU)".

TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/): imported by tests/ and by bench.py's CPU
legs, never by the product package.  The reference runs this path in torch on the host, so the restatement is torch
on the host too (a numpy port would time something the reference does not run); gradients come from autograd, as in
the reference.

Parity is PINNED: tests/golden/ref_cfg1.npz was produced by tests/golden/make_golden_cfg1.py, which imports the
reference's unmodified classes from /root/reference and runs them; tests/test_nerfacto_cpu.py holds this file to
that fixture (outputs, loss and every gradient), and -- in the build container, where /root/reference exists --
to the reference itself at the full 4096 x 48, log2T = 19 size.

What each function follows (reference file:line):
  hash_encode      nerfstudio/field_components/encodings.py:282-349   HashEncoding.hash_fn / pytorch_fwd
  level_scalings   nerfstudio/field_components/encodings.py:252-254   floor(min_res * growth**level)
  sh4              nerfstudio/utils/math.py:27-74                     components_from_spherical_harmonics(levels=4)
  mlp              nerfstudio/field_components/mlp.py:79-97           Linear/ReLU stack, ReLU output activation
  field            nerfstudio/fields/nerfacto_field.py:419-461        TorchNerfactoField.get_density / get_outputs;
                   nerfstudio/field_components/field_heads.py:96-117  softplus density head, sigmoid RGB head
  weights          nerfstudio/cameras/rays.py:155-177                 RaySamples.get_weights
  render           nerfstudio/model_components/renderers.py:97-110 (no background term in this fork), :220, :269-283
"""
import math
import os

import numpy as np
import torch

N_LEVELS = 16
FEATS = 2
MIN_RES, MAX_RES = 16, 2048
PRIMES = (1, 2654435761, 805459861)
# layer shapes (out, in): base 32-64-64-64 (ReLU after every layer), density head 64-1, colour stack 120-32-32
# (ReLU after every layer), RGB head 32-3
LAYERS = ((64, 32), (64, 64), (64, 64), (1, 64), (32, 120), (32, 32), (3, 32))
APPEARANCE_DIM = 40


def level_scalings():
    growth = np.exp((np.log(MAX_RES) - np.log(MIN_RES)) / (N_LEVELS - 1))
    return torch.floor(MIN_RES * growth ** torch.arange(N_LEVELS))


def hash_encode(table, scalings, pos, log2_size):
    """pos [..., 3] -> [..., 32].  Per level: x = pos * scale; cell corners are ceil(x) ("c") and floor(x) ("f") per
    axis; row = ((cx * 1) ^ (cy * 2654435761) ^ (cz * 805459861)) mod 2^log2_size + level * 2^log2_size, evaluated in
    int64; the blend weight of the "c" corner along an axis is x - floor(x)."""
    size = 1 << log2_size
    x = pos[..., None, :] * scalings.view(-1, 1)
    hi = torch.ceil(x).to(torch.int32)
    lo = torch.floor(x).to(torch.int32)
    frac = x - lo
    mult = torch.tensor(PRIMES)
    base = torch.arange(N_LEVELS) * size

    def rows(sel):
        corner = torch.stack([(hi if s else lo)[..., a] for a, s in enumerate(sel)], dim=-1) * mult
        return (corner[..., 0] ^ corner[..., 1] ^ corner[..., 2]) % size + base

    def lerp(a, b, w):
        return a * w + b * (1 - w)

    f = {sel: table[rows(sel)] for sel in [(i, j, k) for i in (0, 1) for j in (0, 1) for k in (0, 1)]}
    wx, wy, wz = frac[..., 0:1], frac[..., 1:2], frac[..., 2:3]
    top = lerp(lerp(f[1, 1, 1], f[0, 1, 1], wx), lerp(f[1, 0, 1], f[0, 0, 1], wx), wy)
    bot = lerp(lerp(f[1, 1, 0], f[0, 1, 0], wx), lerp(f[1, 0, 0], f[0, 0, 0], wx), wy)
    return lerp(top, bot, wz).flatten(-2)


def sh4(d):
    """16 real spherical-harmonic components of a direction, the reference's constants and ordering."""
    x, y, z = d[..., 0], d[..., 1], d[..., 2]
    xx, yy, zz = x ** 2, y ** 2, z ** 2
    c = [torch.full_like(x, 0.28209479177387814),
         0.4886025119029199 * y, 0.4886025119029199 * z, 0.4886025119029199 * x,
         1.0925484305920792 * x * y, 1.0925484305920792 * y * z, 0.9461746957575601 * zz - 0.31539156525251999,
         1.0925484305920792 * x * z, 0.5462742152960396 * (xx - yy),
         0.5900435899266435 * y * (3 * xx - yy), 2.890611442640554 * x * y * z,
         0.4570457994644658 * y * (5 * zz - 1), 0.3731763325901154 * z * (5 * zz - 3),
         0.4570457994644658 * x * (5 * zz - 1), 1.445305721320277 * z * (xx - yy),
         0.5900435899266435 * x * (xx - 3 * yy)]
    return torch.stack(c, dim=-1)


class NerfactoCPU:
    """Parameters as leaf tensors + one fwd/bwd step of configs[0]."""

    def __init__(self, log2_hashmap_size=19, n_images=16, seed=0, params=None):
        self.log2 = int(log2_hashmap_size)
        g = torch.Generator().manual_seed(seed)
        if params is None:
            rows = (1 << self.log2) * N_LEVELS
            params = {"hash_table": (torch.rand(rows, FEATS, generator=g) * 2 - 1) * 1e-3,
                      "embedding": torch.randn(n_images, APPEARANCE_DIM, generator=g)}
            for i, (o, n) in enumerate(LAYERS):                       # torch.nn.Linear default init
                bound = 1 / math.sqrt(n)
                params[f"w{i}"] = (torch.rand(o, n, generator=g) * 2 - 1) * bound
                params[f"b{i}"] = (torch.rand(o, generator=g) * 2 - 1) * bound
        self.p = {k: torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v).clone().requires_grad_(True)
                  for k, v in params.items() if k != "scalings"}
        self.scalings = torch.as_tensor(np.asarray(params["scalings"])) if "scalings" in params else level_scalings()

    def _stack(self, x, first, last):
        for i in range(first, last + 1):
            x = torch.relu(torch.nn.functional.linear(x, self.p[f"w{i}"], self.p[f"b{i}"]))
        return x

    def field(self, pos, dirs, cam):
        """pos [R,S,3], dirs [R,3], cam [R] -> density [R,S,1], rgb [R,S,3] (training mode: embedding looked up)."""
        R, S = pos.shape[:2]
        p = self.p
        base = self._stack(hash_encode(p["hash_table"], self.scalings, pos, self.log2), 0, 2)
        density = torch.nn.functional.softplus(torch.nn.functional.linear(base, p["w3"], p["b3"]))
        with torch.no_grad():
            sh = sh4(dirs[:, None, :].expand(R, S, 3))
        emb = p["embedding"][cam][:, None, :].expand(R, S, APPEARANCE_DIM)
        h = self._stack(torch.cat([sh, base, emb], dim=-1), 4, 5)
        return density, torch.sigmoid(torch.nn.functional.linear(h, p["w6"], p["b6"]))

    @staticmethod
    def weights(delta, density):
        tau = delta * density
        before = torch.cat([torch.zeros_like(tau[..., :1, :]), torch.cumsum(tau[..., :-1, :], dim=-2)], dim=-2)
        return torch.nan_to_num((1 - torch.exp(-tau)) * torch.exp(-before))

    @staticmethod
    def render(w, rgb, starts, ends):
        mid = (starts + ends) / 2
        acc = w.sum(dim=-2)
        depth = torch.clip((w * mid).sum(dim=-2) / (acc + 1e-10), mid.min(), mid.max())
        return (w * rgb).sum(dim=-2), acc, depth

    def forward(self, inp):
        t = {k: torch.as_tensor(v) for k, v in inp.items()}
        density, rgb = self.field(t["pos"], t["dirs"], t["cam"])
        w = self.weights(t["delta"], density)
        out_rgb, acc, depth = self.render(w, rgb, t["starts"], t["ends"])
        loss = torch.nn.functional.mse_loss(out_rgb, t["target"])
        return dict(rgb=out_rgb, accumulation=acc, depth=depth, weights=w, density=density, sample_rgb=rgb, loss=loss)

    def step(self, inp):
        """One forward + backward (no optimizer: configs[0] is "fwd+bwd").  Returns the loss as a float."""
        for v in self.p.values():
            v.grad = None
        out = self.forward(inp)
        out["loss"].backward()
        return float(out["loss"].detach())


def synthetic_inputs(R=4096, S=48, n_images=16, seed=1234):
    """BASELINE.md section 3 shapes: positions U[0,1)^3, one unit direction per ray, deltas U(0,0.05), target U[0,1)."""
    rng = np.random.RandomState(seed)
    pos = rng.uniform(0, 1, size=(R, S, 3)).astype(np.float32)
    d = rng.normal(size=(R, 3))
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    delta = rng.uniform(0, 0.05, size=(R, S, 1)).astype(np.float32)
    ends = np.cumsum(delta, axis=1).astype(np.float32)
    return dict(pos=pos, dirs=d, delta=delta, starts=(ends - delta).astype(np.float32), ends=ends,
                cam=rng.randint(0, n_images, size=(R,)).astype(np.int64),
                target=rng.uniform(0, 1, size=(R, 3)).astype(np.float32))


def time_cfg1(steps=5, warmup=2, R=4096, S=48, log2_hashmap_size=19, threads=None):
    """Median seconds per fwd+bwd of configs[0] with all host threads; returns (rays_per_s, ms, threads)."""
    import time
    threads = threads or os.cpu_count() or 1
    old = torch.get_num_threads()
    torch.set_num_threads(threads)
    try:
        model = NerfactoCPU(log2_hashmap_size=log2_hashmap_size, seed=1234)
        inp = synthetic_inputs(R, S, seed=1234)
        for _ in range(warmup):
            model.step(inp)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            model.step(inp)
            ts.append(time.perf_counter() - t0)
        med = float(np.median(ts))
        return R / med, med * 1e3, torch.get_num_threads()
    finally:
        torch.set_num_threads(old)
