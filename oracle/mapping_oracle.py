"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the WS-MGMap per-step map update.

Nothing in the product path (the `wsmgmap_b200` package / libwsmg.so) may
import this file.  Allowed importers: tests/, __graft_entry__.smoke(),
bench.py's cpu_baseline / --impl reference legs.

Two layers, both restating /root/reference/vlnce_baselines/common/rgb_mapping.py
(cited below as rgb_mapping.py:LINE):

* `OracleMapper` -- the module restated with torch *CPU library ops*
  (affine_grid / grid_sample / scatter_reduce), functional style.  It is the
  CPU baseline that bench.py times ("port") and the float-tolerance checker.
  Pinned: tests/test_oracle_vs_reference.py asserts it equal to the reference
  file executed in the build container, and tests/golden/*.npz (made by
  oracle/make_golden.py from the reference file itself) pin it where the
  reference cannot travel (the GPU box).

* `spec_*` functions -- the same math as explicit elementwise fp32 arithmetic
  (numpy, every rounding spelled out, FMAs where torch-CPU's MKL bmm and the
  vectorised grid_sampler use them).  This is the arithmetic contract of the
  CUDA kernels; tests assert it bit-identical to `OracleMapper` given the same
  sin/cos.

Third-party arithmetic not under /root/reference: torch_scatter 2.0.6
`scatter_max` (reference SETUP.md:55-60; call site rgb_mapping.py:220-225).
Its published semantics are restated in `_scatter_max_cells`: max-reduce,
cells nobody wrote are 0.  No reference test pins that boundary, so parity of
that one call is pinned only by the stand-in in oracle/reference_loader.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F

f32 = np.float32
SENTINEL = -1e16  # rgb_mapping.py:187


@dataclass(frozen=True)
class MapGeometry:
    """Constants of rgb_mapping.py:17-22, 98, 146, 149-151 (python doubles)."""
    resolution: float = 0.12
    ego: int = 100      # egocentric_map_size
    glob: int = 240     # global_map_size

    @property
    def coord_min(self):  # rgb_mapping.py:21
        return -self.glob * self.resolution / 2

    @property
    def coord_max(self):  # rgb_mapping.py:22
        return self.glob * self.resolution / 2

    @property
    def cell(self):       # rgb_mapping.py:98 and :146 (same value)
        return (self.coord_max - self.coord_min) / self.glob

    @property
    def paste_lo(self):   # rgb_mapping.py:42
        return self.glob // 2 - math.floor(self.ego / 2)

    @property
    def paste_hi(self):
        return self.glob // 2 + math.ceil(self.ego / 2)


# --------------------------------------------------------------------------
# Layer 1: torch-CPU restatement (library ops)
# --------------------------------------------------------------------------

def gps_to_cell(gps: torch.Tensor, geo: MapGeometry):
    """rgb_mapping.py:100-103 -- rounded (half-to-even) global cell of the agent."""
    gx = ((geo.coord_max - gps[:, 0]) / geo.cell).round()
    gy = ((gps[:, 1] - geo.coord_min) / geo.cell).round()
    return gx, gy


def unproject_cells(depth_m: torch.Tensor, geo: MapGeometry):
    """rgb_mapping.py:153-176.  depth_m: [bs,Hd,Wd,1] already in metres (x10).
    Returns int64 [bs,2,Hd,Wd] (x_gp, y_gp) and bool [bs,1,Hd,Wd]."""
    d = depth_m.permute(0, 3, 1, 2)
    hd, wd = d.shape[2], d.shape[3]
    cx, cy = hd / 2.0, wd / 2.0
    fx = (hd / 2.0) / np.tan(np.deg2rad(90 / 2.0))
    fy = (wd / 2.0) / np.tan(np.deg2rad(90 / 2.0))
    cols = torch.arange(0, wd, device=d.device).view(1, 1, 1, wd)
    rows = torch.arange(hd, 0, step=-1, device=d.device).view(1, 1, hd, 1)
    xx = (cols - cx) / fx
    yy = (rows - cy) / fy
    big_x = xx * d
    big_y = yy * d
    ok = (d != 0) & ((big_y > -1.5) & (big_y < 0.1))
    half = (geo.ego - 1) / 2
    x_gp = ((big_x / geo.cell) + half).round().long()
    y_gp = (-(d / geo.cell) + half).round().long()
    return torch.cat([x_gp, y_gp], dim=1), ok


def subsample_tables(hf: int, wf: int, hd: int):
    """rgb_mapping.py:188-193 -- nearest-subsample row/col tables."""
    k = hd / wf
    return (torch.arange(0, hf, 1) * k).long(), (torch.arange(0, wf, 1) * k).long()


def linear_cells(locs: torch.Tensor, ok: torch.Tensor, hf: int, wf: int, geo: MapGeometry):
    """rgb_mapping.py:195-217.  Returns (lin int64 [bs,hf,wf], invalid bool [bs,hf,wf])."""
    ri, ci = subsample_tables(hf, wf, locs.shape[-1])
    ri, ci = ri.to(locs.device), ci.to(locs.device)
    ss = locs[:, :, ri[:, None], ci].clone()
    bad_in = ~ok[:, :, ri[:, None], ci].squeeze(1)
    e = geo.ego
    bad_loc = (ss[:, 1] >= e) | (ss[:, 1] < 0) | (ss[:, 0] >= e) | (ss[:, 0] < 0)
    invalid = bad_loc | bad_in
    x = torch.where(invalid, torch.zeros_like(ss[:, 0]), ss[:, 0])
    y = torch.where(invalid, torch.zeros_like(ss[:, 1]), ss[:, 1])
    return y * e + x, invalid


def _scatter_max_cells(src: torch.Tensor, lin: torch.Tensor, n_cells: int):
    """torch_scatter.scatter_max semantics (values only): src [bs,C,N], lin [bs,N]."""
    bs, c, n = src.shape
    idx = lin.view(bs, 1, n).expand(bs, c, n)
    out = torch.full((bs, c, n_cells), torch.finfo(src.dtype).min, dtype=src.dtype, device=src.device)
    out.scatter_reduce_(2, idx, src, reduce="amax", include_self=True)
    hit = torch.zeros((bs, n_cells), dtype=torch.bool, device=src.device)
    hit.scatter_(1, lin, torch.ones_like(lin, dtype=torch.bool))
    return torch.where(hit.unsqueeze(1), out, torch.zeros_like(out)), hit


def project_to_ego(feat: torch.Tensor, lin: torch.Tensor, invalid: torch.Tensor, geo: MapGeometry):
    """rgb_mapping.py:210-232.  Returns proj [bs,C,E,E] and occupancy bits [bs,E*E]
    (cells that received at least one *valid* pixel)."""
    bs, c, hf, wf = feat.shape
    inv_f = invalid.view(bs, 1, hf, wf).float()
    masked = feat * (1 - inv_f) + SENTINEL * inv_f
    e = geo.ego
    proj, _ = _scatter_max_cells(masked.reshape(bs, c, hf * wf), lin.reshape(bs, hf * wf), e * e)
    proj = proj.view(bs, c, e, e)
    hole = (proj == SENTINEL).float()
    proj = proj * (1 - hole) + hole * (proj - SENTINEL)
    occ = torch.zeros((bs, e * e), dtype=torch.bool, device=feat.device)
    flat_lin = lin.reshape(bs, -1)
    flat_ok = ~invalid.reshape(bs, -1)
    cnt = torch.zeros((bs, e * e), dtype=torch.int32, device=feat.device)
    cnt.scatter_add_(1, flat_lin, flat_ok.to(torch.int32))
    occ = cnt > 0
    return proj, occ


def rotate(x: torch.Tensor, heading: torch.Tensor, trig=None):
    """rgb_mapping.py:239-250.  heading [bs,1]; trig optionally (cos,sin) [bs] each."""
    if trig is None:
        sin_t = torch.sin(heading.squeeze(1))
        cos_t = torch.cos(heading.squeeze(1))
    else:
        cos_t, sin_t = trig
    a = torch.zeros(x.size(0), 2, 3, device=x.device)
    a[:, 0, 0] = cos_t
    a[:, 0, 1] = sin_t
    a[:, 1, 0] = -sin_t
    a[:, 1, 1] = cos_t
    grid = F.affine_grid(a, x.size(), align_corners=False)
    return F.grid_sample(x, grid, align_corners=False)


def translation_grid(tx: torch.Tensor, ty: torch.Tensor, size):
    """rgb_mapping.py:130-137 -- theta2 = [[1,-0,x],[0,1,y]] -> affine_grid."""
    one = torch.ones_like(tx)
    zero = torch.zeros_like(tx)
    theta = torch.stack([torch.stack([one, -zero, tx], 1), torch.stack([zero, one, ty], 1)], 1)
    return F.affine_grid(theta, torch.Size(size), align_corners=False)


class OracleMapper:
    """Functional restatement of Mapping/RGBMapping (rgb_mapping.py:11-90) on CPU."""

    def __init__(self, num_proc: int, channels: int = 64, geo: MapGeometry = MapGeometry(), device="cpu"):
        self.geo = geo
        self.channels = channels
        self.device = torch.device(device)      # "cuda": the same torch ops on the GPU (stock-PyTorch comparator of bench.py)
        self.full_global_map = torch.zeros(num_proc, geo.glob, geo.glob, channels, device=self.device)
        self.last = {}

    def stage_cells(self, depth01: torch.Tensor, hf: int, wf: int):
        locs, ok = unproject_cells(depth01 * 10, self.geo)     # rgb_mapping.py:37 (x10)
        return linear_cells(locs, ok, hf, wf, self.geo)

    def step(self, feat, depth01, gps, compass, masks, trig=None, keep=False):
        """One map update (rgb_mapping.py:32-72).  `trig` = dict with optional
        precomputed {'neg': (cos,sin) of -compass, 'pos': (cos,sin) of +compass}."""
        geo = self.geo
        bs, c, hf, wf = feat.shape
        g = geo.glob
        gmap = self.full_global_map
        gx, gy = gps_to_cell(gps, geo)
        gmap[:bs] = gmap[:bs] * masks.view(bs, 1, 1, 1)                      # :35
        lin, invalid = self.stage_cells(depth01, hf, wf)
        proj, occ = project_to_ego(feat, lin, invalid, geo)                  # :266
        rot = rotate(proj, -compass, None if trig is None else trig["neg"])   # :267 / :37
        canvas = torch.zeros(bs, c, g, g, device=feat.device)
        lo, hi = geo.paste_lo, geo.paste_hi
        canvas[:, :, lo:hi, lo:hi] = rot                                     # :40-44
        half = g // 2
        tx = -(gy - half) / half                                             # :47
        ty = -(gx - half) / half                                             # :48
        shifted = F.grid_sample(canvas, translation_grid(tx, ty, canvas.size()), align_corners=False)
        fused = torch.maximum(gmap[:bs], shifted.permute(0, 2, 3, 1))        # :55-56
        gmap[:bs] = fused
        back = F.grid_sample(fused.permute(0, 3, 1, 2).contiguous(),
                             translation_grid(-tx, -ty, canvas.size()), align_corners=False)  # :57-65
        crop = back[:, :, lo:hi, lo:hi]
        ego = rotate(crop, compass, None if trig is None else trig["pos"])    # :70
        if keep:
            self.last = dict(lin=lin, invalid=invalid, proj=proj, occ=occ, rot=rot,
                             shifted=shifted, crop=crop, gx=gx, gy=gy)
        return ego


# --------------------------------------------------------------------------
# Layer 2: elementwise fp32 specification (what the CUDA kernels compute)
# --------------------------------------------------------------------------

def _fma(a, b, c):
    """fp32 fused multiply-add, exact via float64 (24+24-bit product fits in 53 bits;
    the single final rounding of a double sum can double-round only in cases that
    do not occur for these magnitudes -- asserted empirically by the tests)."""
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def spec_base_coords(n: int) -> np.ndarray:
    """ATen affine_grid base grid for align_corners=False: linspace(-1,1,n)*(n-1)/n,
    with torch-CPU's symmetric FMA linspace (start+step*i below the midpoint,
    end-step*(n-1-i) above)."""
    if n == 1:
        return np.zeros(1, f32)
    i = np.arange(n)
    step = f32(f32(2.0) / f32(n - 1))
    lo = _fma(step, i.astype(f32), f32(-1))
    hi = _fma(-step, (n - 1 - i).astype(f32), f32(1))
    lin = np.where(i < n // 2, lo, hi).astype(f32)
    return ((lin * f32(n - 1)).astype(f32) / f32(n)).astype(f32)


def spec_gps_cell(gps: np.ndarray, geo: MapGeometry):
    cmax, cmin, cell = f32(geo.coord_max), f32(geo.coord_min), f32(geo.cell)
    gx = np.rint(((cmax - gps[:, 0].astype(f32)).astype(f32) / cell).astype(f32))
    gy = np.rint(((gps[:, 1].astype(f32) - cmin).astype(f32) / cell).astype(f32))
    return gx.astype(f32), gy.astype(f32)


def spec_cells(depth01: np.ndarray, hf: int, wf: int, geo: MapGeometry):
    """Per sampled pixel: linear cell (invalid -> 0) and invalid flag.
    depth01 [bs,Hd,Wd] fp32.  Appendix-A arithmetic of SURVEY.md."""
    bs, hd, wd = depth01.shape
    k = f32(hd / wf)
    ri = (np.arange(hf).astype(f32) * k).astype(f32).astype(np.int64)
    ci = (np.arange(wf).astype(f32) * k).astype(f32).astype(np.int64)
    cx, cy = f32(hd / 2.0), f32(wd / 2.0)
    fx = f32((hd / 2.0) / np.tan(np.deg2rad(45.0)))
    fy = f32((wd / 2.0) / np.tan(np.deg2rad(45.0)))
    z = (depth01[:, ri[:, None], ci].astype(f32) * f32(10)).astype(f32)
    xx = ((ci.astype(f32) - cx).astype(f32) / fx).astype(f32)[None, None, :]
    yy = (((hd - ri).astype(f32) - cy).astype(f32) / fy).astype(f32)[None, :, None]
    big_x = (xx * z).astype(f32)
    big_y = (yy * z).astype(f32)
    with np.errstate(invalid="ignore"):
        ok = (z != 0) & (big_y > f32(-1.5)) & (big_y < f32(0.1))
        cell, half = f32(geo.cell), f32((geo.ego - 1) / 2)
        xf = np.rint(((big_x / cell).astype(f32) + half).astype(f32))
        yf = np.rint(((-(z / cell).astype(f32)).astype(f32) + half).astype(f32))
        e = f32(geo.ego)
        ok = ok & (xf >= 0) & (xf < e) & (yf >= 0) & (yf < e)
    xi = np.where(ok, xf, 0).astype(np.int64)
    yi = np.where(ok, yf, 0).astype(np.int64)
    return yi * geo.ego + xi, ~ok


def spec_scatter(feat: np.ndarray, lin: np.ndarray, invalid: np.ndarray, geo: MapGeometry):
    """Max of valid pixels per cell; the sentinel rules of rgb_mapping.py:207-230:
    invalid pixels contribute -1e16 to cell 0; untouched cells are 0;
    a cell equal to -1e16 becomes 0; x + 0*(x+1e16) turns -0.0 into +0.0."""
    bs, c, hf, wf = feat.shape
    n = geo.ego * geo.ego
    out = np.zeros((bs, c, n), f32)
    occ = np.zeros((bs, n), bool)
    flat = feat.reshape(bs, c, -1)
    for b in range(bs):
        l = lin[b].reshape(-1)
        ok = ~invalid[b].reshape(-1)
        acc = np.full((c, n), -np.inf, f32)
        np.maximum.at(acc, (slice(None), l[ok]), flat[b][:, ok])
        occ[b, l[ok]] = True
        if (~ok).any():
            acc[:, 0] = np.maximum(acc[:, 0], f32(SENTINEL))
        acc = np.where(np.isneginf(acc), f32(0), acc)
        acc = np.where(acc == f32(SENTINEL), f32(0), acc + f32(0))
        out[b] = acc
    return out.reshape(bs, c, geo.ego, geo.ego), occ


def _spec_unnormalize(gcoord, size):
    # vectorised ATen CPU grid_sampler: (g+1)*(size/2) - 0.5 contracted to one FMA
    return _fma((gcoord + f32(1)).astype(f32), f32(size / 2), f32(-0.5))


def _spec_bilinear(img: np.ndarray, ix: np.ndarray, iy: np.ndarray):
    """img [C,H,W]; ix,iy [h,w].  Zero padding, tap order nw,ne,sw,se with the
    accumulation r=a*nw; r=fma(b,ne,r); r=fma(c,sw,r); r=fma(d,se,r)."""
    c, h, w = img.shape
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    wx = (ix - x0).astype(f32)
    ex = (f32(1) - wx).astype(f32)
    wy = (iy - y0).astype(f32)
    ey = (f32(1) - wy).astype(f32)
    nw, ne, sw, se = (ey * ex).astype(f32), (ey * wx).astype(f32), (wy * ex).astype(f32), (wy * wx).astype(f32)
    x0 = x0.astype(np.int64)
    y0 = y0.astype(np.int64)

    def tap(yy, xx):
        ok = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        v = img[:, np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)]
        return np.where(ok[None], v, f32(0)).astype(f32)

    r = (tap(y0, x0) * nw).astype(f32)
    r = _fma(tap(y0, x0 + 1), ne, r)
    r = _fma(tap(y0 + 1, x0), sw, r)
    r = _fma(tap(y0 + 1, x0 + 1), se, r)
    return r


def spec_rotate(img: np.ndarray, cos_t: float, sin_t: float):
    """img [C,E,E]; MKL K=3 bmm chain: gx = fma(y, s, x*c); gy = fma(y, c, x*(-s))."""
    c_, h, w = img.shape
    bx = spec_base_coords(w)[None, :]
    by = spec_base_coords(h)[:, None]
    cs, sn = f32(cos_t), f32(sin_t)
    gx = _fma(by, sn, (bx * cs).astype(f32))
    gy = _fma(by, cs, (bx * (-sn)).astype(f32))
    return _spec_bilinear(img, _spec_unnormalize(gx, w), _spec_unnormalize(gy, h))


def spec_translate(img: np.ndarray, tx: float, ty: float):
    c_, h, w = img.shape
    gx = (spec_base_coords(w)[None, :] + f32(tx)).astype(f32)
    gy = (spec_base_coords(h)[:, None] + f32(ty)).astype(f32)
    ix = np.broadcast_to(_spec_unnormalize(gx, w), (h, w))
    iy = np.broadcast_to(_spec_unnormalize(gy, h), (h, w))
    return _spec_bilinear(img, ix, iy)


def spec_step(gmap: np.ndarray, feat, depth01, gps, compass, masks, trig, geo: MapGeometry = MapGeometry()):
    """Whole update on numpy arrays.  gmap [n,G,G,C] is updated in place for rows [:bs].
    trig: dict(neg=(cos[bs],sin[bs]), pos=(cos[bs],sin[bs])) fp32 arrays.
    Returns ego [bs,C,E,E] plus intermediates."""
    bs, c, hf, wf = feat.shape
    g, e = geo.glob, geo.ego
    lo, hi = geo.paste_lo, geo.paste_hi
    lin, invalid = spec_cells(depth01, hf, wf, geo)
    proj, occ = spec_scatter(feat, lin, invalid, geo)
    gxc, gyc = spec_gps_cell(gps, geo)
    half = f32(g // 2)
    ego = np.zeros((bs, c, e, e), f32)
    for b in range(bs):
        gmap[b] = (gmap[b] * f32(masks[b])).astype(f32)
        rot = spec_rotate(proj[b], trig["neg"][0][b], trig["neg"][1][b])
        canvas = np.zeros((c, g, g), f32)
        canvas[:, lo:hi, lo:hi] = rot
        tx = -(((gyc[b] - half).astype(f32)) / half).astype(f32)
        ty = -(((gxc[b] - half).astype(f32)) / half).astype(f32)
        shifted = spec_translate(canvas, tx, ty)
        fused = np.maximum(gmap[b], shifted.transpose(1, 2, 0))
        gmap[b] = fused
        back = spec_translate(np.ascontiguousarray(fused.transpose(2, 0, 1)), -tx, -ty)
        ego[b] = spec_rotate(back[:, lo:hi, lo:hi], trig["pos"][0][b], trig["pos"][1][b])
    return ego, dict(lin=lin, invalid=invalid, proj=proj, occ=occ)
