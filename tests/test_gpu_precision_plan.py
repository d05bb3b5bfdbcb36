"""fp16-range guard, per-layer precision plan / auto-tuner, accumulator-truncation constant on non-Gaussian data,
and the reference's standalone helper functions -- all through the C ABI on the GPU."""
import math
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _conv_ref64(x, w4, bias=None):
    """fp64 reference of a 3x3/pad-1 conv: x NHWC fp32, w4 [Cout,Cin,3,3]."""
    y = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w4.double(), None if bias is None else bias.double(), padding=1)
    return y.permute(0, 2, 3, 1)


def _packed(w4, name):
    from marconet_b200.models.networks import _pack_conv_weight
    return _pack_conv_weight(w4, name)


def test_range_guard_flags_overflow_and_reroutes_layer():
    """|x| > 65504 into the default fp16 split: the kernel raises the layer's flag in pinned host memory, poll_range names the
    layer and re-routes it to the bf16 split; the next call is finite and fp32-grade."""
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(2, 16, 16, 128, generator=g) * 3e4).to(dev)          # max ~ 1.3e5 > 65504
    w4 = (torch.randn(128, 128, 3, 3, generator=g) / 34.0).to(dev)
    cw = _packed(w4, "test.range_guard_layer")
    ops.PLAN.pop(cw.name, None)
    cw.precision, cw.x_scale = None, 1.0
    ops.poll_range(dev, reroute=False)
    y = ops.conv2d(x, cw, 3, 3, pad=(1, 1), precision=None)
    torch.cuda.synchronize()
    assert not torch.isfinite(y).all(), "an fp16-split overflow must not produce a plausible finite result"
    with warnings.catch_warnings(record=True) as wrn:
        warnings.simplefilter("always")
        hits = ops.poll_range(dev)
    assert [h.name for h in hits] == [cw.name] and cw.precision == ops.PREC_BF16X3_TC and wrn
    y2 = ops.conv2d(x, cw, 3, 3, pad=(1, 1))
    torch.cuda.synchronize()
    assert ops.poll_range(dev) == []
    ref = _conv_ref64(x, w4)
    rel = ((y2.double() - ref).abs().max() / ref.abs().max()).item()
    print("bf16x3 after reroute: relative max err", rel)
    assert rel < 3e-4
    with pytest.raises(FloatingPointError):           # check_range turns a raised flag into an exception
        cw.precision = ops.PREC_F16X3_TC
        ops.conv2d(x, cw, 3, 3, pad=(1, 1))
        torch.cuda.synchronize()
        ops.check_range(dev)
    ops.PLAN.pop(cw.name, None)


@pytest.mark.parametrize("mag,k", [(3e4, -8), (1.0, 0), (1e-6, 20)])
def test_input_scale_is_exact_and_extends_the_range(mag, k):
    """mn_conv_params.x_scale (power of two) is applied before the split and undone in the epilogue: large inputs no longer
    overflow, tiny inputs no longer fall into fp16's subnormal range, and the result stays fp32-grade."""
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(7)
    x = (torch.randn(2, 16, 16, 128, generator=g) * mag).to(dev)
    w4 = (torch.randn(128, 128, 3, 3, generator=g) / 34.0).to(dev)
    cw = _packed(w4, f"test.xscale_{k}")
    cw.precision, cw.x_scale = ops.PREC_F16X3_TC, 2.0 ** k
    ops.poll_range(dev, reroute=False)
    y = ops.conv2d(x, cw, 3, 3, pad=(1, 1))
    torch.cuda.synchronize()
    assert ops.poll_range(dev, reroute=False) == []
    ref = _conv_ref64(x, w4)
    rel = ((y.double() - ref).abs().max() / ref.abs().max()).item()
    print(f"|x|~{mag:g}, x_scale 2^{k}: relative max err {rel:.2e}")
    assert rel < 2e-5
    ops.PLAN.pop(cw.name, None)


@pytest.mark.parametrize("kind", ["positive", "negative_weights", "lognormal"])
def test_accumulator_truncation_fix_on_one_signed_data(kind):
    """The epilogue undoes the expected truncation shrink of the main accumulator with an empirical constant (conv_tc2.cu `dfix`,
    1.5e-8 per accumulation step, fitted on Gaussian data).  One-signed / heavy-tailed operands make every partial sum grow
    monotonically -- the worst case for a truncating accumulator: the result must stay fp32-grade and essentially unbiased."""
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(11)
    x = torch.randn(2, 16, 16, 512, generator=g)
    w4 = torch.randn(128, 512, 3, 3, generator=g) / 68.0
    if kind == "positive":
        x, w4 = x.abs() + 0.5, w4.abs()
    elif kind == "negative_weights":
        x, w4 = x.abs() + 0.5, -w4.abs()
    else:
        x, w4 = torch.exp(x * 1.5), w4.abs() * torch.exp(torch.randn(128, 512, 3, 3, generator=g))
    x, w4 = x.to(dev), w4.to(dev)
    cw = _packed(w4, "test.dfix_" + kind)
    ref = _conv_ref64(x, w4)
    for prec, tol_max, tol_bias in ((ops.PREC_F16X3_TC, 2e-5, 6e-6), (ops.PREC_FP32_SIMT, 2e-5, 6e-6)):
        y = ops.conv2d(x, cw, 3, 3, pad=(1, 1), precision=prec)
        inner = (slice(None), slice(1, -1), slice(1, -1))                 # interior pixels: all 9 taps, K = 4608
        relerr = ((y.double() - ref) / ref.abs().clamp_min(1e-30))[inner]
        mx, bias = relerr.abs().max().item(), relerr.mean().item()
        print(f"{kind} precision {prec}: max rel err {mx:.2e}, mean signed rel err {bias:+.2e}")
        assert mx < tol_max and abs(bias) < tol_bias
    ops.PLAN.pop(cw.name, None)


def test_tune_precision_on_out_of_range_checkpoint(checkpoints):
    """A checkpoint whose activations leave fp16's range (ResNet stem x1000: the BN-free ReLU network is positively homogeneous, so
    every feature map grows 1000x, |x| ~ 1e6): the untuned default raises the range flag; pipeline.tune_precision installs
    per-layer input scales / formats; the tuned run is finite, unflagged and matches the oracle on the SAME checkpoint."""
    from marconet_b200 import ops, pipeline
    from marconet_b200.models import networks
    from oracle import restate, synth
    dev = torch.device("cuda:0")
    sd = {k: v.clone() for k, v in checkpoints["encoder"].items()}
    sd["resnet.conv1.weight"] = sd["resnet.conv1.weight"] * 1000.0
    saved_plan = dict(ops.PLAN)
    try:
        ops.PLAN.clear()
        nets = {}
        for key, cls, s in (("tspgan", networks.TSPGAN, checkpoints["tspgan"]), ("encoder", networks.TextContextEncoderV2, sd),
                            ("sr", networks.TSPSRNet, checkpoints["sr"])):
            m = cls()
            m.load_state_dict(s, strict=True)
            nets[key] = m.eval().to(dev)
        lq = synth.make_lq(1, 0)
        labels, locs = [synth.make_labels(3, 0)], synth.make_locs(1, 3)
        ops.poll_range(dev, reroute=False)
        nets["encoder"](lq.to(dev))
        torch.cuda.synchronize()
        flagged = ops.poll_range(dev, reroute=False)
        assert flagged and all(c.name.startswith("encoder.resnet") for c in flagged), [c.name for c in flagged]
        report = pipeline.tune_precision(nets["encoder"], nets["tspgan"], nets["sr"], lq.to(dev), labels, locs.to(dev))
        scaled = [r for r in report if r["name"].startswith("encoder.resnet") and r["x_scale"] < 1.0]
        print("tuned layers:", len(report), "with x_scale < 1:", len(scaled), "formats:",
              {p: sum(1 for r in report if r["precision"] == p) for p in (0, 1, 2)})
        assert scaled, "the out-of-range ResNet layers must receive a down-scale"
        logits, _, w = nets["encoder"](lq.to(dev))
        torch.cuda.synchronize()
        assert ops.poll_range(dev, reroute=False) == [] and torch.isfinite(logits).all()
        ol, _, ow = restate.encoder_forward(sd, lq)
        rel_l = ((logits.cpu() - ol).abs().max() / ol.abs().max()).item()
        rel_w = ((w.cpu() - ow).abs().max() / ow.abs().max()).item()
        print("tuned encoder vs oracle (x1000 checkpoint): relative max err logits", rel_l, "w", rel_w)
        assert rel_l < 1e-3 and rel_w < 1e-3
        # the plan survives a re-pack (keyed by layer name) and a JSON round trip
        import json, os, tempfile
        with tempfile.TemporaryDirectory() as d:
            pipeline.save_precision_plan(os.path.join(d, "plan.json"))
            plan = json.load(open(os.path.join(d, "plan.json")))
        assert any(v[1] < 1.0 for v in plan.values())
        nets["encoder"]._invalidate()
        logits2, _, _ = nets["encoder"](lq.to(dev))
        assert torch.equal(logits2, logits)
    finally:
        ops.PLAN.clear()
        ops.PLAN.update(saved_plan)


def test_tune_precision_on_regular_checkpoint_keeps_parity(gpu_models, checkpoints):
    """On the regular synthetic checkpoints the tuner must leave parity where it was (golden config 2, 1e-3)."""
    import os
    from marconet_b200 import ops, pipeline
    from oracle.make_golden import STRIDES, case_inputs
    dev = torch.device("cuda:0")
    saved_plan = dict(ops.PLAN)
    try:
        inp = case_inputs("config2")
        report = pipeline.tune_precision(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"], inp["lq"].to(dev), inp["labels"],
                                         inp["locs"].to(dev))
        worst = max(min(r["err_f16x3"], r["err_bf16x3"]) for r in report)
        print("layers", len(report), "worst best-format relative layer error", worst,
              "formats", {p: sum(1 for r in report if r["precision"] == p) for p in (0, 1, 2)})
        out = pipeline.restore_lines(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"], inp["lq"].to(dev), inp["labels"], inp["locs"].to(dev))
        g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config2.npz"))
        err = float(np.abs(out["sr"].float().cpu().contiguous().reshape(-1)[::STRIDES["sr"]].numpy() - g["sr"]).max())
        print("tuned plan: sr max-abs err vs reference golden", err)
        assert err <= 1e-3
    finally:
        for cw in pipeline.conv_layers(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"]):
            cw.precision, cw.x_scale = None, 1.0
        ops.PLAN.clear()
        ops.PLAN.update(saved_plan)


def test_reference_helper_functions():
    """swish / calc_mean_std_4D / adaptive_instance_normalization (reference networks.py:492-493, 518-533) as standalone functions."""
    from marconet_b200.models import networks
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    a = torch.randn(2, 7, 5, 9, generator=g) * 2 + 0.3
    b = torch.randn(2, 7, 4, 6, generator=g) * 0.5 - 1.0
    s = networks.swish(a.to(dev)).cpu()
    assert torch.allclose(s, a * torch.sigmoid(a), atol=1e-6)
    m, sd = networks.calc_mean_std_4D(a.to(dev))
    var = a.view(2, 7, -1).var(dim=2) + 1e-5
    assert torch.allclose(m.cpu().view(2, 7), a.view(2, 7, -1).mean(2), atol=1e-6) and torch.allclose(sd.cpu().view(2, 7), var.sqrt(), atol=1e-6)
    out = networks.adaptive_instance_normalization(a.to(dev), b.to(dev)).cpu()
    bm, bv = b.view(2, 7, -1).mean(2).view(2, 7, 1, 1), (b.view(2, 7, -1).var(2) + 1e-5).sqrt().view(2, 7, 1, 1)
    am, av = a.view(2, 7, -1).mean(2).view(2, 7, 1, 1), var.sqrt().view(2, 7, 1, 1)
    assert torch.allclose(out, (a - am) / av * bv + bm, atol=1e-5)


@pytest.mark.parametrize("masked", [False, True])
def test_groupnorm_statistics_accumulated_in_the_conv_epilogue(masked):
    """mn_conv_params.gn_stats_out: the producing conv's epilogue accumulates sum / sum of squares per (sample, 32-channel group) of
    its output; finalised statistics must equal the separate read pass (mn_groupnorm_stats) over the stored tensor."""
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(17)
    n, h, w, cin, cout = 5, 32, 32, 128, 256
    x = torch.randn(n, h, w, cin, generator=g).to(dev)
    w4 = (torch.randn(cout, cin, 3, 3, generator=g) / 34.0).to(dev)
    bias = torch.randn(cout, generator=g).to(dev)
    cw = _packed(w4, f"test.gn_stats_{masked}")
    vw = torch.tensor([32, 7, 19, 32, 1], dtype=torch.int32, device=dev) if masked else None
    y, mr = ops.conv2d(x, cw, 3, 3, pad=(1, 1), bias=bias, act=ops.ACT_LRELU02, valid_w=vw, gn_stats=True)
    ref = ops.groupnorm_stats(y, valid_w=vw)
    torch.cuda.synchronize()
    assert mr.shape == ref.shape == (n, cout // 32, 2)
    err_mean = (mr[..., 0] - ref[..., 0]).abs().max().item()
    err_rstd = ((mr[..., 1] - ref[..., 1]).abs() / ref[..., 1].abs()).max().item()
    print("epilogue GN statistics vs read pass: mean abs err", err_mean, "rstd rel err", err_rstd)
    assert err_mean < 1e-5 and err_rstd < 1e-5
    # and the exact fp32 kernel (no fused statistics there) takes the separate pass transparently
    y0, mr0 = ops.conv2d(x, cw, 3, 3, pad=(1, 1), bias=bias, act=ops.ACT_LRELU02, valid_w=vw, gn_stats=True, precision=ops.PREC_FP32_SIMT)
    assert (mr0[..., 0] - ref[..., 0]).abs().max().item() < 1e-3
    ops.PLAN.pop(cw.name, None)
