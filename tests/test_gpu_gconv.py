"""Kernel-level GPU parity of the tcgen05 grouped convolution kernels, called through the C ABI (ops.gconv_raw /
ops.gconv_wgrad_raw), against the oracle's MP_Conv arithmetic (oracle.mp_weight + the asymmetric 'same' pad +
conv2d of oracle.mp_conv, models/model_internals.py:253-271) evaluated in float64 on the same bf16-rounded operands.

  forward / data gradient (gconv2): rel-L2 <= 4e-3  (one bf16 rounding of the output)
  weight gradient (gwgrad2, fp32 accumulators):                    rel-L2 <= 1e-3
  W-PREP operand (bf16, tap-major) vs bf16(oracle.mp_weight):      <= 1 bf16 ulp per element

Shapes: every (Cin, Cout, H) combination of SURVEY Appendix D plus the 64x64 level of config C, mixed kernel sizes
per group (1 / 3 / 5 / 7), an expert that receives no rows, fewer live rows than the launch capacity, the ones-channel
input (Cin 33 padded to 64), and the fused epilogue flags."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2
from oracle import hdmoe_oracle as O

pytestmark = pytest.mark.gpu

FWD_TOL = 4e-3
WGRAD_TOL = 1e-3


def _round_up(x, k):
    return (x + k - 1) // k * k


def _cin_pad(cin):
    return _round_up(cin, 32) if cin % 32 == 0 else _round_up(cin, 64)


def _ref_conv(x64, w64):
    """conv2d with the reference's asymmetric 'same' padding (oracle.mp_conv, models/model_internals.py:261-271)."""
    k = w64.shape[-1]
    lo = (k - 1) // 2
    return F.conv2d(F.pad(x64, (lo, k - 1 - lo, lo, k - 1 - lo)), w64)


class _Case:
    """One grouped layer: raw MP_Conv weights per expert -> W-PREP (ops.WeightPrep, both operand layouts) -> rows."""

    def __init__(self, R, H, W, cin, cout, ks, counts, gain=1.0, seed=0):
        from hdmoe_b200 import ops
        self.R, self.H, self.W, self.cin, self.cout, self.ks, self.counts = R, H, W, cin, cout, list(ks), list(counts)
        gen = torch.Generator().manual_seed(seed + R * 131 + H + cin * 7 + cout)
        self.cin_pad = _cin_pad(cin)
        self.cin_rows = cin if cin % 32 == 0 else (cin // 32) * 32
        E = len(ks)
        row_e = sum(([e] * c for e, c in enumerate(counts)), [])
        self.n_rows = len(row_e)
        assert self.n_rows <= R
        self.row_e_host = row_e + [-1] * (R - self.n_rows)
        self.row_e = torch.tensor(self.row_e_host, dtype=torch.int32, device="cuda")
        self.n_rows_dev = torch.tensor([self.n_rows], dtype=torch.int32, device="cuda")
        self.w_raw = [torch.randn(cout, cin, k, k, generator=gen) * (0.5 + e) for e, k in enumerate(ks)]
        self.gain = gain
        self.wrow, self.wrow_t = [], []
        r = rt = 0
        for k in ks:
            self.wrow.append(r)
            self.wrow_t.append(rt)
            r += k * k * cout
            rt += k * k * self.cin_rows
        self.rows_total, self.rows_total_t = r, rt
        self.w_fwd = torch.zeros(r, self.cin_pad, dtype=torch.bfloat16, device="cuda")
        self.w_bwd = torch.zeros(max(rt, 1), cout, dtype=torch.bfloat16, device="cuda")
        self.w_dev = [w.cuda().contiguous() for w in self.w_raw]
        ent = []
        for e, k in enumerate(ks):
            d = dict(w=self.w_dev[e], gain=gain, layout="taps", cin_pad=self.cin_pad,
                     out=self.w_fwd[self.wrow[e]:self.wrow[e] + k * k * cout])
            if self.cin_rows > 0:
                d.update(layout2="taps_t", cin_rows=self.cin_rows, cout_pad=cout,
                         out2=self.w_bwd[self.wrow_t[e]:self.wrow_t[e] + k * k * self.cin_rows])
            ent.append(d)
        ops.WeightPrep(ent, torch.device("cuda")).run(force=False)
        torch.cuda.synchronize()
        # oracle weights, rounded to bf16 like the operand
        self.w_hat = [O.mp_weight(w, gain) for w in self.w_raw]
        self.w_hat_bf = [w.to(torch.bfloat16) for w in self.w_hat]
        x = torch.randn(R, cin, H, W, generator=gen)
        self.x_bf = x.to(torch.bfloat16)
        xn = torch.zeros(R, H, W, self.cin_pad, dtype=torch.bfloat16)
        xn[..., :cin] = self.x_bf.permute(0, 2, 3, 1)
        self.x_nhwc = xn.cuda()
        self.gen = gen

    def operand_matches_oracle(self):
        """the tap-major bf16 operand W-PREP wrote == bf16(oracle.mp_weight), element by element (<= 1 ulp)"""
        got = self.w_fwd.float().cpu()
        for e, k in enumerate(self.ks):
            blk = got[self.wrow[e]:self.wrow[e] + k * k * self.cout].view(k * k, self.cout, self.cin_pad)
            ref = self.w_hat[e].permute(2, 3, 0, 1).reshape(k * k, self.cout, self.cin)
            assert float(blk[..., self.cin:].abs().max()) == 0 if self.cin_pad > self.cin else True
            ulp = ref.abs().clamp_min(1e-30) * 2.0 ** -7
            assert bool(((blk[..., :self.cin] - ref).abs() <= ulp).all()), f"expert {e}"

    def ref_forward(self, scale=None, act=0, res=None, res_a=0.0, res_b=1.0):
        out = torch.zeros(self.n_rows, self.H, self.W, self.cout, dtype=torch.float64)
        for r in range(self.n_rows):
            e = self.row_e_host[r]
            y = _ref_conv(self.x_bf[r:r + 1].double(), self.w_hat_bf[e].double())[0].permute(1, 2, 0)
            if scale is not None:
                y = y * scale[r].double()
            if act == 1:
                y = O.mp_silu(y)
            if res is not None:
                y = res_a * res[r].double() + res_b * y.to(torch.bfloat16).double()
            out[r] = y
        return out


SHAPES = [
    # R, H, W, cin, cout, ks, counts
    (2, 16, 16, 64, 64, [1], [2]),
    (3, 32, 32, 32, 32, [3], [3]),
    (6, 32, 32, 32, 32, [3, 3, 5, 5], [1, 2, 0, 2]),          # an expert without rows, n_rows < cap
    (5, 16, 16, 128, 64, [3, 5], [2, 3]),
    (5, 16, 16, 96, 64, [3, 5], [2, 3]),
    (4, 32, 32, 96, 32, [5, 3], [2, 2]),
    (4, 32, 32, 33, 32, [3, 5], [2, 1]),                       # ones-channel input: Cin 33 padded to 64
    (4, 32, 32, 64, 32, [3, 5, 1], [1, 2, 1]),
    (4, 32, 32, 64, 64, [3, 5], [2, 2]),
    (4, 16, 16, 32, 64, [1, 1], [3, 1]),                       # 1x1 skip
    (3, 64, 64, 32, 32, [3, 5], [1, 2]),                       # config C level
    (3, 64, 64, 64, 64, [5, 3], [2, 1]),
    (3, 32, 32, 128, 64, [1, 3], [1, 2]),
    (3, 32, 32, 32, 64, [3], [3]),                             # router trunk conv 1 (dense, one group)
    (3, 32, 32, 64, 128, [3], [2]),                            # router trunk conv 2
    (2, 32, 32, 128, 128, [3], [2]),                           # router trunk conv 3
    (3, 64, 64, 32, 128, [3], [3]),
    (5, 20, 12, 32, 96, [7, 1, 3], [2, 1, 2]),                 # ragged spatial size, k = 7
    (40, 16, 16, 64, 64, [3, 3, 5, 5], [5, 10, 12, 13]),
    (70, 32, 32, 64, 64, [3, 3, 5, 5], [9, 12, 20, 25]),       # more tiles than one wave of the scheduler per SM pair
]


def _ids(s):
    return f"R{s[0]}_{s[1]}x{s[2]}_ci{s[3]}_co{s[4]}_k{''.join(map(str, s[5]))}"


@pytest.mark.parametrize("shape", SHAPES, ids=_ids)
def test_gconv_forward_vs_oracle(shape):
    from hdmoe_b200 import ops
    R, H, W, cin, cout, ks, counts = shape
    c = _Case(*shape)
    c.operand_matches_oracle()
    y = ops.gconv_raw(c.x_nhwc, c.w_fwd, cout, c.rows_total, c.row_e, c.n_rows_dev, ks, c.wrow)
    torch.cuda.synchronize()
    ref = c.ref_forward()
    got = y.float().cpu()[:c.n_rows].double()
    assert torch.isfinite(got).all()
    assert rel_l2(got, ref) < FWD_TOL
    # per-row: no row may hide behind the others (a wrong expert / kernel size on one row is a routing bug)
    for r in range(c.n_rows):
        assert rel_l2(got[r], ref[r]) < 2 * FWD_TOL, (r, c.row_e_host[r])
    if c.n_rows < R:
        assert float(y[c.n_rows:].float().abs().max()) == 0.0     # unused tail rows are zero-filled


@pytest.mark.parametrize("flags", [(True, 1, False), (False, 0, True), (True, 1, True), (True, 0, False)],
                         ids=["scale_silu", "residual", "scale_silu_residual", "scale"])
@pytest.mark.parametrize("shape", [(4, 32, 32, 64, 64, [3, 5], [2, 2]), (6, 32, 32, 32, 32, [3, 3, 5, 5], [1, 2, 0, 2]),
                                   (9, 16, 16, 64, 64, [3, 5], [4, 5])], ids=_ids)
def test_gconv_fused_epilogue_vs_oracle(shape, flags):
    """out = res_a * residual + res_b * mp_silu(scale * conv): the eval-mode fusion of Unet_block
    (models/model_components.py:240-253)."""
    from hdmoe_b200 import ops
    R, H, W, cin, cout, ks, counts = shape
    use_scale, act, use_res = flags
    c = _Case(*shape, seed=5)
    scale = (torch.rand(R, cout, generator=c.gen) + 0.5) if use_scale else None
    res = torch.randn(R, H, W, cout, generator=c.gen).to(torch.bfloat16) if use_res else None
    t = 0.3
    cc = math.sqrt((1 - t) ** 2 + t ** 2)
    ra, rb = ((1 - t) / cc, t / cc) if use_res else (0.0, 1.0)
    y = ops.gconv_raw(c.x_nhwc, c.w_fwd, cout, c.rows_total, c.row_e, c.n_rows_dev, ks, c.wrow,
                      scale=None if scale is None else scale.cuda(), act=act,
                      residual=None if res is None else res.cuda(), res_a=ra, res_b=rb)
    torch.cuda.synchronize()
    ref = c.ref_forward(scale=scale, act=act, res=res, res_a=ra, res_b=rb)
    assert rel_l2(y.float().cpu()[:c.n_rows].double(), ref) < FWD_TOL


@pytest.mark.parametrize("shape", [s for s in SHAPES if s[3] >= 32], ids=_ids)
def test_gconv_data_gradient_vs_oracle(shape):
    """dX = conv_transpose(dY, W_hat) through the same kernel with the transposed, tap-flipped operand W-PREP writes
    (HDMOE_WLAYOUT_TAPS_T), against torch.nn.grad.conv2d_input in float64."""
    from hdmoe_b200 import ops
    R, H, W, cin, cout, ks, counts = shape
    c = _Case(*shape, seed=9)
    if cout % 32 != 0:
        pytest.skip("data-gradient K must be a multiple of 32")
    dy = torch.randn(R, H, W, cout, generator=c.gen).to(torch.bfloat16)
    dx = ops.gconv_raw(dy.cuda(), c.w_bwd, c.cin_rows, c.rows_total_t, c.row_e, c.n_rows_dev, ks, c.wrow_t)
    torch.cuda.synchronize()
    got = dx.float().cpu().double()
    for r in range(c.n_rows):
        e = c.row_e_host[r]
        k = ks[e]
        ref = torch.nn.grad.conv2d_input((1, cin, H, W), c.w_hat_bf[e].double(), dy[r:r + 1].double().permute(0, 3, 1, 2),
                                         padding=(k - 1) // 2)[0].permute(1, 2, 0)[..., :c.cin_rows]
        assert rel_l2(got[r], ref) < 2 * FWD_TOL, (r, e)


# weight-gradient-only shapes: wide inputs with a 7x7 expert (one kernel row per unit: the largest unit tables)
WGRAD_EXTRA = [(2, 8, 8, 256, 64, [7, 3], [1, 1]), (2, 16, 16, 192, 64, [7], [2]), (3, 16, 16, 256, 32, [5, 7], [2, 1])]


@pytest.mark.parametrize("shape", [s for s in SHAPES if s[4] in (32, 64, 128) and s[1] % 4 == 0] + WGRAD_EXTRA, ids=_ids)
def test_gwgrad_vs_oracle(shape):
    """dW_hat[tap][o][c] = sum_{rows of e, pixels} dY[r, q, o] * Xpad[r, q + delta_tap, c] (fp32 accumulators) against
    torch.nn.grad.conv2d_weight in float64 on the same bf16 operands; experts without rows keep a zero block."""
    from hdmoe_b200 import ops
    R, H, W, cin, cout, ks, counts = shape
    c = _Case(*shape, seed=11)
    dy = torch.randn(R, H, W, cout, generator=c.gen).to(torch.bfloat16)
    dw = torch.zeros(c.rows_total, c.cin_pad, dtype=torch.float32, device="cuda")
    ops.gconv_wgrad_raw(c.x_nhwc, dy.cuda(), dw, c.row_e, c.n_rows_dev, ks, c.wrow)
    torch.cuda.synchronize()
    got = dw.cpu().double()
    lo = 0
    for e, n in enumerate(counts):
        k = ks[e]
        blk = got[c.wrow[e]:c.wrow[e] + k * k * cout].view(k * k, cout, c.cin_pad)
        if n == 0:
            assert float(blk.abs().max()) == 0.0, e
            continue
        xin = c.x_bf[lo:lo + n].double()
        g = dy[lo:lo + n].double().permute(0, 3, 1, 2)
        ref = torch.nn.grad.conv2d_weight(xin, (cout, cin, k, k), g, padding=(k - 1) // 2)
        ref = ref.permute(2, 3, 0, 1).reshape(k * k, cout, cin)
        assert rel_l2(blk[..., :cin], ref) < WGRAD_TOL, e
        lo += n
    # a second accumulation adds (the kernel accumulates into dW; the caller zeroes it once per step)
    ops.gconv_wgrad_raw(c.x_nhwc, dy.cuda(), dw, c.row_e, c.n_rows_dev, ks, c.wrow)
    torch.cuda.synchronize()
    assert rel_l2(dw.cpu().double(), 2 * got) < 1e-6
