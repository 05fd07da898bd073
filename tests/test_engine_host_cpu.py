"""Launch sequencing of the product engines checked WITHOUT a GPU: `UNetEngine.forward`, `Rollout.denoise` and
`VAEDecoderEngine.decode` run on CPU tensors with every C-ABI call served by the host emulator of include/wiw.h
(tests/helpers/cabi_emulator.py -- test infrastructure, fp32 arithmetic) and are compared with the golden outputs of the
reference's vendored fork.  What this pins on the CPU: the order and operands of ~750 launches per forward (residual /
AlphaBlender wiring, the collapsed cross-attention vectors, the fused time-embedding GEMM and its column offsets, the frame-position
tables, skip concatenation, up / down-sampling), the sampler loop, and the decode tail.  The kernels' own arithmetic is what
the `-m gpu` tests check."""
import os
import sys

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import ops
from wiw_b200 import weights as W
from wiw_b200.config import TINY_UNET, TINY_VAE
from wiw_b200.engine import UNetEngine, VAEDecoderEngine, sinusoid
from wiw_b200.pipeline import Rollout
from wiw_b200.scheduler import EulerSchedule
from oracle import svd_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
import cabi_emulator  # noqa: E402


@pytest.fixture()
def emu(monkeypatch):
    return cabi_emulator.install(monkeypatch)


def rel_rms(a, b):
    return float(((a - b).pow(2).mean() / b.pow(2).mean()).sqrt())


def test_unet_forward_sequencing_vs_golden(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_tiny.npz"))
    sample, enc, aids = torch.from_numpy(g["sample"]), torch.from_numpy(g["enc"]), torch.from_numpy(g["added_action_ids"])
    nb, F, _, h, w = sample.shape
    unet = UNetEngine(TINY_UNET, W.random_init(W.unet_param_spec(TINY_UNET), 11), "cpu")
    unet.set_condition(enc[1], 0.02, aids[0])
    x_in = ops.nchw_to_nhwc(sample.flatten(0, 1).half().contiguous(), 64)
    t_emb = sinusoid(torch.tensor([float(g["timestep"])]), TINY_UNET.block_out_channels[0]).half()
    taps = {}
    v = unet.forward(x_in, t_emb, taps=taps)
    out = ops.nhwc_to_nchw(v, 4).float().view(nb, F, 4, h, w)
    for k in ("down.0.0", "mid", "up.3.2"):
        got = ops.nhwc_to_nchw(taps[k], taps[k].shape[-1]).float()
        assert rel_rms(got, torch.from_numpy(g["tap." + k])) < 1e-2, k
    assert rel_rms(out, torch.from_numpy(g["out"])) < 1e-2
    n_gemm = len(emu.calls)
    assert n_gemm > 250                                    # every Linear / conv of the forward went through the descriptor path


def test_rollout_and_decode_sequencing_vs_reference_pipeline(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "rollout_tiny.npz"))
    n = int(g["num_steps"])
    unet = UNetEngine(TINY_UNET, W.random_init(W.unet_param_spec(TINY_UNET), 11), "cpu")
    vae = VAEDecoderEngine(TINY_VAE, W.random_init(W.vae_decoder_param_spec(TINY_VAE), 12), "cpu")
    ro = Rollout(unet, vae, "cpu", use_cuda_graph=False)
    lat0 = (torch.from_numpy(g["latents_raw"]).half() * EulerSchedule(n).init_noise_sigma)[0]
    lat = ro.denoise(lat0, torch.from_numpy(g["image_latents"])[1], torch.from_numpy(g["image_embeddings"])[1],
                     torch.from_numpy(g["added_action_ids"])[0], num_steps=n)
    final = ops.nhwc_to_nchw(lat, 4).float()
    ref_lat = torch.from_numpy(g["final_latents"])[0]
    rng = float(ref_lat.max() - ref_lat.min())
    psnr_lat = 20 * np.log10(rng / float((final - ref_lat).pow(2).mean().sqrt()))
    assert psnr_lat > 45.0, psnr_lat
    u8, f32 = ro.decode(lat, want_f32=True)
    frames = (f32 / 2 + 0.5).clamp(0, 1).numpy()
    assert O.psnr(frames, g["frames"][0]) >= 45.0
    assert O.psnr(u8.numpy().astype(np.float32) / 255.0, g["frames"][0]) >= 45.0


def test_vae_encoder_sequencing_vs_golden(emu, golden_dir):
    """the SD-VAE encoder + quant_conv on the implicit-GEMM / GroupNorm wrappers reproduces the
    reference's fp32 posterior mean (frontend_tiny.npz) -- asymmetric stride-2 taps, shortcut convs, single-head attention and
    the quant_conv fold."""
    from wiw_b200.engine import VAEEncoderEngine
    from wiw_b200.frontend import vae_encoder_param_spec
    g = np.load(os.path.join(golden_dir, "frontend_tiny.npz"))
    c = TINY_VAE
    esd = W.random_init(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), 13)
    enc = VAEEncoderEngine(c, esd, "cpu")
    got = enc.encode_mode(torch.from_numpy(g["image"])).float()
    ref = torch.from_numpy(g["latent_mode"])
    assert got.shape == ref.shape
    assert rel_rms(got, ref) < 1e-2, rel_rms(got, ref)


def test_worker_request_path_on_host(emu, tmp_path):
    """Request dict -> response dict through the real WorldModelWorker / Navigator / Rollout / resizer on CPU tensors (C ABI
    emulated): shapes, dtypes, determinism of the seeded generator, and the resize stage being byte-identical to the reference's
    own post-processing (PIL BICUBIC + to_tensor + (clip*255).astype(uint8), api_models/__init__.py:159-166,229)."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    from wiw_b200 import wm_worker as WW
    from wiw_b200.frontend import vae_encoder_param_spec
    args = WW.build_parser().parse_args(["--width", "128", "--height", "64", "--out_width", "48", "--out_height", "40", "--num_frames", "8",
                                         "--action_input_channel", "8", "--num_inference_steps", "2", "--device", "cpu"])
    unet = UNetEngine(TINY_UNET, W.random_init(W.unet_param_spec(TINY_UNET), 11), "cpu")
    vae = VAEDecoderEngine(TINY_VAE, W.random_init(W.vae_decoder_param_spec(TINY_VAE), 12), "cpu")
    c = TINY_VAE
    esd = W.random_init(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), 13)
    torch.manual_seed(3)
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2,
                                                          patch_size=32, projection_dim=TINY_UNET.cross_attention_dim)).eval()
    nav = WW.Navigator(args, Rollout(unet, vae, "cpu", use_cuda_graph=False), clip, esd, "cpu")
    worker = WW.WorldModelWorker(args, nav)
    rng = np.random.default_rng(0)
    req = {"b_action": np.array([[4, 1, 2, 2, 1, 3, 1, 1]], dtype=np.int64), "b_image": rng.integers(0, 256, size=(1, 4, 64, 128), dtype=np.uint8),
           "save_dirs": [str(tmp_path / "a")], "request_model_name": "igen", "return_objects": [True]}
    nav.generator.manual_seed(1)
    out = worker.do_some_tasks(req)
    WW.check_outputdict(out)
    pf = out["pred_frames"]
    assert pf.shape == (1, 8, 3, 40, 48) and pf.dtype == np.uint8 and pf.std() > 1.0
    nav.generator.manual_seed(1)
    full = nav.inference(WW.action_ids_idx_encode(req["b_action"]), req["b_image"])          # uint8 [1, 8, 3, 64, 128]
    ref = (np.clip(WW.images_to_tensor(full, save_size=(48, 40)).numpy(), 0, 1) * 255).astype(np.uint8)
    assert np.array_equal(pf, ref)


def test_files_mode_writes_reference_bytes(emu, tmp_path):
    """Files mode (<save_dir>/<idx>.jpg + action_seq.json): the JPEGs written from the device-side resize equal the ones the
    reference's images_to_tensor -> save_predict chain writes from the same full-resolution frames."""
    from wiw_b200 import wm_worker as WW

    class Nav:                                             # fixed full-resolution frames: isolates the output stage
        device = torch.device("cpu")

        def inference(self, actions, images_u8, ow=None, oh=None, on_device=False, **kw):
            g = torch.Generator().manual_seed(5)
            return torch.randint(0, 256, (actions.shape[0], 8, 3, 64, 128), generator=g, dtype=torch.uint8)

    args = WW.build_parser().parse_args(["--width", "128", "--height", "64", "--out_width", "48", "--out_height", "40", "--num_frames", "8"])
    worker = WW.WorldModelWorker(args, Nav())
    acts = np.array([[4, 1, 2, 2, 1, 3, 1, 1]], dtype=np.int64)
    req = {"b_action": acts, "b_image": np.zeros((1, 3, 64, 128), np.uint8), "save_dirs": [str(tmp_path / "new")], "request_model_name": "igen"}
    out = worker.do_some_tasks(req)
    assert out == {"save_dirs": req["save_dirs"]}
    full = Nav().inference(torch.zeros(1, 8, 8), None).numpy()
    WW.save_predict(WW.images_to_tensor(full, save_size=(48, 40)).numpy(), acts, [str(tmp_path / "ref")])
    names = sorted(os.listdir(tmp_path / "ref"))
    assert names == sorted(os.listdir(tmp_path / "new")) and len(names) == 9
    for n in names:
        assert open(tmp_path / "ref" / n, "rb").read() == open(tmp_path / "new" / n, "rb").read(), n
