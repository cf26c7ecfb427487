"""GPU unit tests of the GEMM kernels through the C-ABI test hook: the tcgen05/TMEM/TMA kernel and the
CUDA-core kernel against a torch fp32 reference on the same bf16-rounded operands (so the only
difference is fp32 accumulation order: tolerance 2e-5 of the row-wise scale)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda:0"
SHAPES = [  # (M, N, K): model shapes, ragged M tails, N not a multiple of the tile, long K
    (2560, 1536, 384), (2560, 384, 512), (2560, 2304, 384), (2560, 384, 1152), (2560, 400, 384), (2560, 512, 384),
    (1, 384, 384), (10, 1536, 384), (127, 400, 384), (129, 64, 64), (300, 1024, 384), (190 * 7, 2304, 384),
    (1000, 2052, 2048), (4096, 256, 128), (486400 // 8, 1536, 384), (32768, 384, 1152), (20000, 1024, 384),
]


@pytest.mark.parametrize("path", [0, 1, 2, 3, 4, 5])  # 4 / 5: persistent kernel (BN = 256 or 192 / BN = 192)
def test_gemm_paths_match_fp32_reference(engine_bf16, report, path):
    g = torch.Generator().manual_seed(path)
    worst = 0.0
    for (M, N, K) in SHAPES:
        A = (torch.randn(M, K, generator=g) * 2).to(torch.bfloat16)
        W = (torch.randn(N, K, generator=g) * 0.1).to(torch.bfloat16)
        ref = (A.float().to(DEV) @ W.float().to(DEV).T)
        out = engine_bf16.debug_gemm_bf16(A.to(DEV), W.to(DEV), path)
        scale = float(ref.abs().max())
        err = float((out - ref).abs().max()) / scale
        worst = max(worst, err)
        assert err <= 2e-5, f"path {path} shape {(M, N, K)}: normalised error {err}"
    report(test="gemm_bf16", path=path, worst_norm_err=worst)


def test_bf16_model_same_tokens_with_and_without_tensor_cores(engine_bf16, report):
    """The two GEMM back-ends see identical bf16 operands; logits agree to accumulation-order noise."""
    from music2midi_b200 import synthetic as syn

    wave = syn.audio_noise(4, 77).to(DEV)
    cond = torch.zeros(4, 2, dtype=torch.long, device=DEV)
    mel = engine_bf16.logmel(wave)
    emb = engine_bf16.condition(mel, cond)
    engine_bf16.set_flags(no_tensor_cores=False)
    toks, lg = engine_bf16.generate_from_embeds(emb, 64, return_logits=True)
    forced = torch.zeros(4, 64, dtype=torch.long, device=DEV)
    forced[:, : toks.shape[1]] = toks
    engine_bf16.set_flags(no_tensor_cores=True)
    _, lg2 = engine_bf16.generate_from_embeds(emb, 64, forced=forced, return_logits=True)
    engine_bf16.set_flags(no_tensor_cores=False)
    d = float((lg - lg2).abs().max())
    report(test="bf16_tc_vs_simt_logits", max_abs=d)
    assert d <= 0.1


def test_two_contexts_on_two_devices_in_one_process(state_dict):
    """The dynamic-shared-memory opt-in of the tcgen05 kernels is per device: a second context on another GPU of the
    same process must run them too (skipped on single-GPU boxes)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from music2midi_b200 import synthetic as syn
    from music2midi_b200.engine import Engine

    wave = syn.audio_noise(3, 5)
    cond = torch.zeros(3, 2, dtype=torch.long)
    outs = []
    for d in (0, 1):
        eng = Engine(torch.device("cuda", d), "bf16")
        eng.load_state_dict(state_dict)
        outs.append(eng.generate(wave.to(f"cuda:{d}"), cond.to(f"cuda:{d}"), 24).cpu())
        eng.close()
    assert torch.equal(outs[0], outs[1])


def test_fp32_split_product_gemms_match_cuda_core_gemms(engine_fp32, report):
    """fp32 parity mode runs its GEMMs as three-term bf16 split products on tcgen05 (six products, fp32 accumulate);
    against the plain FFMA kernels on the same weights the logits agree to fp32 rounding noise and the greedy tokens
    are identical."""
    from music2midi_b200 import synthetic as syn

    wave = syn.audio_noise(4, 78).to(DEV)
    cond = torch.zeros(4, 2, dtype=torch.long, device=DEV)
    emb = engine_fp32.condition(engine_fp32.logmel(wave), cond)
    toks, lg = engine_fp32.generate_from_embeds(emb, 64, return_logits=True)
    engine_fp32.set_flags(no_f32_tc=True)
    try:
        toks2, lg2 = engine_fp32.generate_from_embeds(emb, 64, return_logits=True)
    finally:
        engine_fp32.set_flags(no_f32_tc=False)
    assert torch.equal(toks, toks2)
    d = float((lg - lg2).abs().max())
    report(test="fp32_tc_vs_simt_logits", max_abs=d)
    assert d <= 5e-4
