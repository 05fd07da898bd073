"""CPU-only checks of the host side: the C-ABI library loads and exports every symbol
include/wiw.h declares, the product scheduler is bit-exact against the golden tables, the
weight inventory matches the reference shapes, and the product refuses to run without a GPU."""
import os
import re

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import lib as L
from wiw_b200 import weights as W
from wiw_b200.config import FULL_UNET, FULL_VAE, TINY_UNET
from wiw_b200.pipeline import sample_latent_noise_
from wiw_b200.scheduler import EulerSchedule
from oracle import svd_oracle as O

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(REPO, "include", "wiw.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(wiw_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = L.load()
    names = _declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/wiw.h but not exported"
        assert n in L.SYMBOLS, f"{n} has no ctypes signature in wiw_b200/lib.py"
    assert lib.wiw_abi_version() == L.ABI_VERSION
    assert isinstance(lib.wiw_last_error(), bytes)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(L.WiwError):
        L.require_device()
    from wiw_b200 import ops
    with pytest.raises(L.WiwError):
        ops.layer_norm(torch.zeros(4, 64, dtype=torch.float16), ops.NormWeight(torch.ones(64), torch.zeros(64)))


@pytest.mark.parametrize("n", [10, 25, 30])
def test_product_scheduler_bit_exact(golden_dir, n):
    g = np.load(os.path.join(golden_dir, "scheduler.npz"))
    s = EulerSchedule(n)
    assert s.sigmas.numpy().tobytes() == g[f"sigmas_{n}"].tobytes()
    assert s.timesteps.numpy().tobytes() == g[f"timesteps_{n}"].tobytes()
    assert np.float32(s.init_noise_sigma.item()).tobytes() == g[f"init_noise_sigma_{n}"].tobytes()
    assert [s.index_for_timestep(t) for t in s.timesteps] == list(g[f"index_{n}"])
    # the per-step scalars are the oracle's 0-dim tensor expressions
    o = O.EulerSchedule(n)
    for i, sc in enumerate(s.all_steps()):
        sig = o.sigmas[i]
        assert sc.div == float((sig ** 2 + 1) ** 0.5) and sc.c_out == float(-sig / (sig ** 2 + 1) ** 0.5)
        assert sc.dt == float(o.sigmas[i + 1] - sig) and sc.sig2p1 == float(sig ** 2 + 1)


def test_param_inventory():
    spec = W.unet_param_spec(FULL_UNET)
    assert sum(int(np.prod(s)) for _, s in spec) == 1_528_603_626
    names = [n for n, _ in spec]
    assert len(set(names)) == len(names)
    assert "down_blocks.0.attentions.0.temporal_transformer_blocks.0.ff_in.net.0.proj.weight" in names
    assert dict(spec)["add_action_proj.proj.weight"] == (256, 168)
    vspec = W.vae_decoder_param_spec(FULL_VAE)
    assert dict(vspec)["time_conv_out.weight"] == (3, 3, 3, 1, 1)
    a = W.random_init(W.unet_param_spec(TINY_UNET)[:6], 3)
    b = W.random_init(W.unet_param_spec(TINY_UNET)[:6], 3)
    assert all(torch.equal(a[k], b[k]) for k in a)
    with pytest.raises(KeyError):
        W.check_state_dict({}, spec)


def test_latent_noise_matches_oracle():
    g = torch.Generator().manual_seed(0)
    noise = torch.randn(8, 4, 2, 32, generator=g)
    acts = torch.tensor([4, 2, 2, 1, 3, 3, 0, 2])
    ref = O.sample_latent_noise(noise[None].clone(), acts[None])[0]
    assert torch.equal(sample_latent_noise_(noise.clone(), acts), ref)
    with pytest.raises(AssertionError):
        sample_latent_noise_(torch.zeros(2, 4, 2, 8), torch.tensor([4, 2]))


def test_tile_shape_pickers():
    """Host-side M-tile / N-tile selection (ops._box3, _tbox, pick_bn): every pick is a legal tile (product 128, power-of-two
    extents the per-warp TMA store boxes can split) and fills the north-star feature maps as DESIGN.md section 4 states."""
    from wiw_b200 import ops

    def util3(w, h, n, box):
        bx, by, bz = box
        return (min(w, bx) / bx if w < bx else w / (-(-w // bx) * bx)) * (h / (-(-h // by) * by)) * (n / (-(-n // bz) * bz))

    for (w, h, n) in [(128, 72, 28), (64, 36, 28), (32, 18, 28), (16, 9, 28), (16, 8, 16), (8, 4, 16), (2, 1, 16), (1024, 576, 14), (16, 9, 3)]:
        bx, by, bz = ops._box3(w, h, n)
        assert bx * by * bz == 128 and all(v & (v - 1) == 0 for v in (bx, by, bz))
    assert ops._box3(128, 72, 28) == (128, 1, 1) and ops._box3(32, 18, 28) == (32, 2, 2) and ops._box3(16, 9, 28) == (16, 2, 4)
    assert util3(16, 9, 28, ops._box3(16, 9, 28)) == pytest.approx(0.9) and util3(16, 9, 28, (16, 8, 1)) == pytest.approx(0.5625)
    for (s, f) in [(9216, 14), (2304, 14), (576, 14), (144, 14), (128, 8), (2, 8)]:
        bx, by = ops._tbox(s, f)
        assert bx * by == 128
    assert ops._tbox(144, 14) == (16, 8) and ops._tbox(576, 14) == (64, 2) and ops._tbox(9216, 14) == (128, 1)
    # N-tiles: fewest tiles, then narrowest; GEGLU tiles divide N
    assert ops.pick_bn(320) == 160 and ops.pick_bn(640) == 224 and ops.pick_bn(960) == 256 and ops.pick_bn(1280) == 256
    assert ops.pick_bn(2560, geglu=True) == 256 and ops.pick_bn(512, geglu=True) == 256 and ops.pick_bn(64) == 64
    with pytest.raises(ValueError):
        ops.pick_bn(100)


def test_pair_layout_for_cfg_split():
    from wiw_b200 import sharding
    pairs, owner = sharding.pair_layout(3, 8)
    assert pairs == 4 and owner == [(0, 0), (2, 0), (4, 0)]
    assert sharding.pair_branch_indices(3, 0, 8) == [0] and sharding.pair_branch_indices(3, 1, 8) == [0]
    assert sharding.pair_branch_indices(3, 5, 8) == [2] and sharding.pair_branch_indices(3, 7, 8) == []
    pairs, owner = sharding.pair_layout(5, 4)
    assert owner == [(0, 0), (2, 0), (0, 1), (2, 1), (0, 2)] and sharding.pair_branch_indices(5, 3, 4) == [1, 3]


def test_library_contains_native_tcgen05_and_tma_code():
    """The built library is native sm_100a code: its SASS holds the tensor-core (UTCHMMA, incl. the 2-CTA form), TMEM (LDTM / STTM)
    and TMA (UTMALDG / UTMASTG) instructions the hot kernels are written around (mnemonics: B200_PROFILING.md)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", L.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in sass or "SM100" in sass.upper()
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG.5D", "UTMALDG.5D.2CTA", "UTMASTG.4D", "UTCBAR"):
        assert mnemonic in sass, f"{mnemonic} missing from the library's SASS"
    for kernel in ("conv_gemm_kernel", "attn_spatial_kernel", "temporal_attn_mma_kernel", "layernorm_lpr_kernel", "gn_apply_kernel"):
        assert kernel in sass


def test_gemm_register_budgets_survive_the_compiler():
    """conv_gemm_kernel rebalances registers by role with setmaxnreg (producers 40, epilogue 232).  A function CALL anywhere in the
    kernel (a printf, a noinline helper) makes ptxas drop the per-role budgets and allocate everything under the smallest one:
    nothing fails, the epilogue just spills ~1 KB per thread and the K <= 640 layers run up to 2x slower.
    Guard: the lean kernels keep their USETMAXREG pair, contain no call and no local-memory traffic, and use the packed f32x2
    forms the issue-bound epilogue is written around."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", L.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    kernels, name = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name is not None and "/*" in line:
            kernels[name].append(line)
    lean = {k: v for k, v in kernels.items() if re.search(r"conv_gemm_kernelILi[12]ELi[02]ELi1E", k)}     # <CG, GEGLU, LEAN = 1>
    assert len(lean) == 4, sorted(kernels)[:8]
    for k, body in lean.items():
        text = "\n".join(body)
        assert "USETMAXREG.DEALLOC" in text and "USETMAXREG.TRY_ALLOC" in text, f"{k}: setmaxnreg pair missing"
        assert not re.search(r"\b(STL|LDL)\b", text), f"{k}: local-memory spills in a lean GEMM kernel"
        assert not re.search(r"\bCALL\.ABS|\bCALL\.REL(?!\.NOINC)", text), f"{k}: a real function call voids the setmaxnreg budgets"
        assert re.search(r"\bFFMA2\b|\bFADD2\b", text), f"{k}: packed f32x2 arithmetic missing"
    attn = [v for k, v in kernels.items() if "attn_spatial_kernel" in k]
    assert attn and all(re.search(r"\bFMNMX3\b", "\n".join(v)) and re.search(r"\bFFMA2\b", "\n".join(v)) for v in attn)
