"""TEST INFRASTRUCTURE ONLY: a host emulator of the C ABI declared in include/wiw.h.

Each method executes the *documented meaning* of one entry point on host memory with numpy / torch fp32, so the product's
host logic -- descriptor building in `ops.py`, launch sequencing in `engine.py` / `pipeline.py` -- can be checked against the
golden vectors on a box without a GPU.  It is injected by `install(monkeypatch)` inside tests and never imported by the
product (`world-in-world_b200/` fails loudly without its CUDA library); it says nothing about the kernels themselves, which
the `-m gpu` tests cover."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F


def _arr(ptr, n, dtype):
    if not ptr:
        return None
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class FakeLib:
    """Executes a wiw_gemm_desc the way include/wiw.h documents it (fp32 accumulation, fp16 output)."""

    def __init__(self):
        self.calls = []

    def wiw_conv_gemm(self, d, stream):
        ad, ast = list(d.a_dim), list(d.a_stride)
        od, ost = list(d.out_dim), list(d.out_stride)
        taps, cpt = d.num_taps, d.chunks_per_tap
        cin = cpt * 64
        assert d.k_total == taps * cin and d.bx * d.by * max(d.bz, 1) == 128 and d.bn % 32 == 0 and 32 <= d.bn <= 256
        a_span = 1 + sum((ad[i] - 1) * ast[i] for i in range(5))
        A = _arr(d.a, a_span, np.float16).astype(np.float32)
        Wm = _arr(d.w, d.n_total * d.k_total, np.float16).astype(np.float32).reshape(d.n_total, d.k_total)
        n_out = d.n_total // 2 if d.geglu else d.n_total
        assert od[0] == n_out
        X, Y, N = od[1], od[2], od[3]
        acc = np.zeros((N, Y, X, d.n_total), np.float32)
        xs, ys, ns = np.arange(X), np.arange(Y), np.arange(N)
        for t in range(taps):
            cb, ox, p, oy = d.tap_cbase[t], d.tap_ox[t], d.tap_p[t], d.tap_oy[t]
            xi, yi = xs + ox, ys + oy
            ok = ((xi >= 0) & (xi < ad[1]))[None, None, :] & ((yi >= 0) & (yi < ad[3]))[None, :, None] & np.ones((N, 1, 1), bool)
            base = (ns[:, None, None] * ast[4] + np.clip(yi, 0, ad[3] - 1)[None, :, None] * ast[3] + p * ast[2]
                    + np.clip(xi, 0, ad[1] - 1)[None, None, :] * ast[1])
            ch = cb + np.arange(cin)
            chok = ch < ad[0]                                                  # channels past the tensor read as zero (TMA OOB fill)
            patch = A[(base[..., None] + np.clip(ch, 0, ad[0] - 1) * ast[0])] * chok * ok[..., None]
            acc += patch @ Wm[:, t * cin:(t + 1) * cin].T
        if d.bias:
            acc += _arr(d.bias, d.n_total, np.float32)
        if d.tab:
            grow = (ns[:, None, None] * Y + ys[None, :, None]) * X + xs[None, None, :]
            rows = (grow // d.tab_div) % d.tab_mod
            stride = d.tab_stride if d.tab_stride > 0 else d.n_total
            tab = _arr(d.tab, (d.tab_mod - 1) * stride + d.n_total, np.float16).astype(np.float32)
            acc += tab[rows[..., None] * stride + np.arange(d.n_total)]
        if d.act == 1:
            acc = acc / (1.0 + np.exp(-acc))
        elif d.act == 2:
            acc = F.gelu(torch.from_numpy(acc)).numpy()
        elif d.act == 3:
            acc = acc / (1.0 + np.exp(-1.702 * acc))
        if d.geglu:
            bn, hb = d.bn, d.bn // 2
            v = acc.reshape(N, Y, X, d.n_total // bn, bn)
            gate = torch.from_numpy(np.ascontiguousarray(v[..., hb:]))
            acc = (v[..., :hb] * F.gelu(gate).numpy()).reshape(N, Y, X, n_out)
        out = acc * d.s_acc
        o_span = 1 + sum((od[i] - 1) * ost[i] for i in range(4))
        idx = (ns[:, None, None, None] * ost[3] + ys[None, :, None, None] * ost[2] + xs[None, None, :, None] * ost[1]
               + np.arange(n_out)[None, None, None, :] * ost[0])
        if d.res:
            out = out + d.s_res * _arr(d.res, o_span, np.float16).astype(np.float32)[idx]
        if d.aux:
            out = out + d.s_aux * _arr(d.aux, o_span, np.float16).astype(np.float32)[idx]
        O = _arr(d.out, o_span, np.float16)
        O[idx] = out.astype(np.float16)
        self.calls.append((d.bx, d.by, d.bz, d.bn))
        return 0

    def wiw_last_error(self):
        return b""

    # ---- small helpers
    @staticmethod
    def _t(ptr, shape, dtype=np.float16):
        n = int(np.prod(shape))
        return _arr(ptr if isinstance(ptr, int) else getattr(ptr, "value", ptr), n, dtype).reshape(shape)

    def wiw_groupnorm_workspace_floats(self, samples, groups):
        return samples * 1025 * groups * 2

    def wiw_groupnorm(self, x, y, samples, rows, C, G, gamma, beta, eps, silu, ws, stream):
        xv = torch.from_numpy(self._t(x, (samples, rows, C)).astype(np.float32))
        g, b = torch.from_numpy(self._t(gamma, (C,), np.float32).copy()), torch.from_numpy(self._t(beta, (C,), np.float32).copy())
        r = F.group_norm(xv.permute(0, 2, 1), G, g, b, eps).permute(0, 2, 1)
        if silu:
            r = F.silu(r)
        self._t(y, (samples, rows, C))[...] = r.numpy().astype(np.float16)
        return 0

    def wiw_layernorm(self, x, y, M, C, gamma, beta, eps, tab, tab_div, tab_mod, stream):
        xv = torch.from_numpy(self._t(x, (M, C)).copy())
        if tab:
            rows = (torch.arange(M) // tab_div) % tab_mod
            xv = xv + torch.from_numpy(self._t(tab, (tab_mod, C)).copy())[rows]          # fp16 add, as the reference does
        g, b = torch.from_numpy(self._t(gamma, (C,), np.float32).copy()), torch.from_numpy(self._t(beta, (C,), np.float32).copy())
        self._t(y, (M, C))[...] = F.layer_norm(xv.float(), (C,), g, b, eps).numpy().astype(np.float16)
        return 0

    def wiw_attention_spatial(self, qkv, out, images, S, C, stream):
        t = torch.from_numpy(self._t(qkv, (images, S, 3, C // 64, 64)).astype(np.float32))
        q, k, v = (t[:, :, i].permute(0, 2, 1, 3) for i in range(3))                       # [img, heads, S, 64]
        o = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(images, S, C)
        self._t(out, (images, S, C))[...] = o.numpy().astype(np.float16)
        return 0

    def wiw_attention_temporal(self, qkv, out, B, Fr, S, C, stream):
        t = torch.from_numpy(self._t(qkv, (B, Fr, S, 3, C // 64, 64)).astype(np.float32))
        q, k, v = (t[:, :, :, i].permute(0, 2, 3, 1, 4) for i in range(3))                 # [B, S, heads, F, 64]
        o = F.scaled_dot_product_attention(q, k, v).permute(0, 3, 1, 2, 4).reshape(B, Fr, S, C)
        self._t(out, (B, Fr, S, C))[...] = o.numpy().astype(np.float16)
        return 0

    def wiw_emb_silu(self, te, act, noise, out, R, C, stream):
        e = torch.from_numpy(self._t(te, (C,)).copy())[None] + torch.from_numpy(self._t(act, (R, C)).copy())
        e = e + torch.from_numpy(self._t(noise, (R, C)).copy())                            # two fp16 adds
        self._t(out, (R, C))[...] = F.silu(e.float()).numpy().astype(np.float16)
        return 0

    def wiw_concat_channels(self, a, b, out, M, C1, C2, stream):
        o = self._t(out, (M, C1 + C2))
        o[:, :C1] = self._t(a, (M, C1))
        o[:, C1:] = self._t(b, (M, C2))
        return 0

    def wiw_upsample_nearest2x(self, x, y, N, H, W, C, stream):
        xv = self._t(x, (N, H, W, C))
        self._t(y, (N, 2 * H, 2 * W, C))[...] = xv.repeat(2, axis=1).repeat(2, axis=2)
        return 0

    def wiw_softmax_rows(self, s, rows, S, scale, stream):
        v = self._t(s, (rows, S))
        v[...] = torch.softmax(torch.from_numpy(v.astype(np.float32)) * scale, dim=-1).numpy().astype(np.float16)
        return 0

    def wiw_transpose(self, x, y, batch, R, C, stream):
        self._t(y, (batch, C, R))[...] = self._t(x, (batch, R, C)).transpose(0, 2, 1)
        return 0

    def wiw_nchw_to_nhwc(self, src, dst, N, C, P, cs, stream):
        d = self._t(dst, (N, P, cs))
        d[...] = 0
        d[:, :, :C] = self._t(src, (N, C, P)).transpose(0, 2, 1)
        return 0

    def wiw_nhwc_to_nchw(self, src, dst, N, C, P, cs, stream):
        self._t(dst, (N, C, P))[...] = self._t(src, (N, P, cs))[:, :, :C].transpose(0, 2, 1)
        return 0

    def wiw_sampler_prepare(self, lat, img, xin, Fr, hw, div, stream):
        l = self._t(lat, (Fr, hw, 4)).astype(np.float32)
        x = self._t(xin, (2, Fr, hw, 64))
        scaled = (l / np.float32(div)).astype(np.float16)
        x[:, :, :, 0:4] = scaled[None]
        x[0, :, :, 4:8] = 0
        x[1, :, :, 4:8] = self._t(img, (hw, 4))[None]
        return 0

    def wiw_sampler_step(self, v, lat, gs, Fr, hw, vstride, c_out, sig2p1, sigma, dt, stream):
        vv = self._t(v, (2, Fr, hw, vstride))[..., :4]
        vu, vc = torch.from_numpy(vv[0].copy()), torch.from_numpy(vv[1].copy())
        g = torch.from_numpy(self._t(gs, (Fr,)).copy())[:, None, None]
        comb = vu + g * (vc - vu)                                                          # fp16 tensor ops (pipeline :606-608)
        x = torch.from_numpy(self._t(lat, (Fr, hw, 4)).copy()).float()
        pred = (comb * torch.tensor(c_out, dtype=torch.float32)).half().float() + x / np.float32(sig2p1)
        deriv = (x - pred) / np.float32(sigma)
        self._t(lat, (Fr, hw, 4))[...] = (x + deriv * np.float32(dt)).numpy().astype(np.float16)
        return 0

    def wiw_vae_finish(self, x, w, bias, out_u8, out_f32, Fr, HW, cs, stream):
        xv = torch.from_numpy(self._t(x, (Fr, HW, cs))[..., :3].astype(np.float32))       # [F, HW, 3]
        wt = torch.from_numpy(self._t(w, (3, 3, 3), np.float32).copy())                    # (co, ci, kt)
        b = torch.from_numpy(self._t(bias, (3,), np.float32).copy())
        y = F.conv1d(xv.permute(1, 2, 0), wt, b, padding=1).permute(2, 1, 0)              # [HW, 3, F] -> [F, 3, HW]
        if out_f32:
            self._t(out_f32, (Fr, 3, HW), np.float32)[...] = y.numpy()
        if out_u8:
            self._t(out_u8, (Fr, 3, HW), np.uint8)[...] = ((y / 2 + 0.5).clamp(0, 1) * 255).round().numpy().astype(np.uint8)
        return 0


    def wiw_resize_bicubic_u8(self, src, tmp, dst, planes, H, W, oh, ow, bounds_h, kk_h, ksize_h, bounds_v, kk_v, ksize_v, lut, stream):
        """Pillow's integer two-pass resampler from the host-built tables (22-bit fixed point, 8-bit intermediate), then the LUT."""
        x = self._t(src, (planes, H, W), np.uint8).astype(np.int64)
        bh, kh = self._t(bounds_h, (ow, 2), np.int32), self._t(kk_h, (ow, ksize_h), np.int32).astype(np.int64)
        bv, kv = self._t(bounds_v, (oh, 2), np.int32), self._t(kk_v, (oh, ksize_v), np.int32).astype(np.int64)
        t = np.zeros((planes, H, ow), np.int64)
        for xx in range(ow):
            x0, n = int(bh[xx, 0]), int(bh[xx, 1])
            t[:, :, xx] = np.clip(((x[:, :, x0:x0 + n] * kh[xx, :n]).sum(-1) + (1 << 21)) >> 22, 0, 255)
        self._t(tmp, (planes, H, ow), np.uint8)[...] = t.astype(np.uint8)
        o = np.zeros((planes, oh, ow), np.int64)
        for yy in range(oh):
            y0, n = int(bv[yy, 0]), int(bv[yy, 1])
            o[:, yy, :] = np.clip(((t[:, y0:y0 + n, :] * kv[yy, :n, None]).sum(1) + (1 << 21)) >> 22, 0, 255)
        self._t(dst, (planes, oh, ow), np.uint8)[...] = self._t(lut, (256,), np.uint8)[o]
        return 0


    def wiw_resize_separable_f32(self, src, tmp, dst, planes, channels, H, W, oh, ow, bounds_h, w_h, ksize_h, bounds_v, w_v, ksize_v,
                                 mean, inv_std, stream):
        """Banded separable resampler of uint8 planes (value / 255) + (x - mean[c]) * inv_std[c], fp16 out."""
        x = self._t(src, (planes, H, W), np.uint8).astype(np.float64) / 255.0
        bh, kh = self._t(bounds_h, (ow, 2), np.int32), self._t(w_h, (ow, ksize_h), np.float32).astype(np.float64)
        bv, kv = self._t(bounds_v, (oh, 2), np.int32), self._t(w_v, (oh, ksize_v), np.float32).astype(np.float64)
        t = np.zeros((planes, H, ow))
        for xx in range(ow):
            x0, n = int(bh[xx, 0]), int(bh[xx, 1])
            t[:, :, xx] = (x[:, :, x0:x0 + n] * kh[xx, :n]).sum(-1)
        o = np.zeros((planes, oh, ow))
        for yy in range(oh):
            y0, n = int(bv[yy, 0]), int(bv[yy, 1])
            o[:, yy, :] = (t[:, y0:y0 + n, :] * kv[yy, :n, None]).sum(1)
        m, s = self._t(mean, (channels,), np.float32), self._t(inv_std, (channels,), np.float32)
        c = np.arange(planes) % channels
        self._t(dst, (planes, oh, ow), np.float16)[...] = ((o - m[c][:, None, None]) * s[c][:, None, None]).astype(np.float16)
        return 0

    def wiw_patchify(self, x, out, B, C, G, ps, kpad, stream):
        img = torch.from_numpy(self._t(x, (B, C, G * ps, G * ps), np.float16).astype(np.float32))
        cols = F.unfold(img, kernel_size=ps, stride=ps).transpose(1, 2).reshape(B * G * G, C * ps * ps)      # (c, ky, kx)-major columns
        o = self._t(out, (B * G * G, kpad), np.float16)
        o[...] = 0
        o[:, : C * ps * ps] = cols.numpy().astype(np.float16)
        return 0

    def wiw_attention_small(self, qkv, out, batch, S, heads, d, scale, stream):
        C = heads * d
        t = torch.from_numpy(self._t(qkv, (batch, S, 3, heads, d), np.float16).astype(np.float32))
        q, k, v = (t[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        o = torch.softmax(q @ k.transpose(-1, -2) * scale, dim=-1) @ v
        self._t(out, (batch, S, C), np.float16)[...] = o.permute(0, 2, 1, 3).reshape(batch, S, C).numpy().astype(np.float16)
        return 0


def install(monkeypatch):
    """Route wiw_b200.lib / ops to a FakeLib for the duration of one test (CPU tensors, no streams)."""
    from wiw_b200 import lib as L
    from wiw_b200 import ops
    lib = FakeLib()
    monkeypatch.setattr(L, "require_device", lambda: lib)
    monkeypatch.setattr(L, "stream_ptr", lambda: None)
    monkeypatch.setattr(ops, "_chk", lambda x: None)

    class _Stream:
        cuda_stream = 0
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    return lib
