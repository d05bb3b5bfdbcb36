"""Self-contained flow helpers: host-side integer logic on CPU, full flow on GPU vs the oracle."""
import io

import pytest
import torch


def test_decode_labels_matches_oracle_clear_labels():
    from marconet_b200.pipeline import decode_labels
    from oracle import restate
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(64, 6736, generator=g)
    logits[5] = logits[4]            # a repeat
    logits[7, 6735] = 100.0          # a blank
    assert decode_labels(logits) == restate.clear_labels(logits)


def test_lr_to_center_halfwidth_matches_training_loop():
    from marconet_b200.pipeline import lr_to_center_halfwidth
    g = torch.Generator().manual_seed(1)
    lr = torch.rand(3, 32, generator=g)
    ref = lr.clone()
    for b in range(lr.size(0)):                      # Train/tspgan/models/tspgan_model.py:332-336, verbatim semantics
        for n in range(0, lr.size(1), 2):
            ref[b][n] = (lr[b][n + 1] + lr[b][n]) / 2.0
            ref[b][n + 1] = (lr[b][n + 1] - lr[b][n]) / 2.0
    assert torch.equal(lr_to_center_halfwidth(lr), ref)


def test_load_checkpoint_variants(checkpoints):
    from marconet_b200.models import networks
    from marconet_b200.pipeline import load_checkpoint
    sd = checkpoints["tspgan"]
    m = networks.TSPGAN()
    load_checkpoint(m, {"params": sd})
    ema = {k: v + 1 for k, v in sd.items()}
    load_checkpoint(m, {"params": sd, "params_ema": ema})
    assert torch.equal(m.state_dict()["TextGenerator.conv1.bias"], ema["TextGenerator.conv1.bias"])
    load_checkpoint(m, {"module." + k: v for k, v in sd.items()})
    buf = io.BytesIO()
    torch.save({"params": sd}, buf)
    buf.seek(0)
    load_checkpoint(m, buf)
    assert torch.equal(m.state_dict()["TextGenerator.conv1.bias"], sd["TextGenerator.conv1.bias"])
    with pytest.raises(RuntimeError):
        load_checkpoint(m, {"params": {k: v for k, v in list(sd.items())[:-1]}})


@pytest.mark.gpu
def test_restore_lines_self_contained_flow(gpu_models, checkpoints):
    """Encoder-predicted labels + boxes drive TSPGAN and TSPSRNet; compared with the same flow through the oracle."""
    from marconet_b200.pipeline import decode_labels, lr_to_center_halfwidth, restore_lines
    from marconet_b200.testing import synth
    from oracle import restate
    dev = torch.device("cuda:0")
    lq = synth.make_lq(1, 31)
    out = restore_lines(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"], lq.to(dev), max_chars=3)
    ol, olr, ow = restate.encoder_forward(checkpoints["encoder"], lq)
    labels = [torch.tensor(decode_labels(ol[0])[:3], dtype=torch.long).reshape(-1, 1)]
    assert [l.tolist() for l in out["labels"]] == [l.tolist() for l in labels], "decoded labels must be bit-exact"
    locs = lr_to_center_halfwidth(olr)
    oi, o64, o32 = restate.tspgan_forward(checkpoints["tspgan"], ow.repeat(labels[0].shape[0], 1), labels[0])
    assert (out["prior"][0].cpu() - oi).abs().max().item() <= 1e-3
    # the boxes are float predictions: only when both sides truncate to the same window integers is the SR image comparable
    from marconet_b200.models.networks import char_windows
    n = labels[0].shape[0]
    same = all(char_windows(out["locs"].cpu(), [n], w_, h_)[0] == char_windows(locs, [n], w_, h_)[0] for w_, h_ in ((512, 16), (1024, 32)))
    assert (out["locs"].cpu() - locs).abs().max().item() <= 1e-4
    if same:
        osr = restate.tspsr_forward(checkpoints["sr"], lq, [o64], [o32], locs)
        assert (out["sr"].cpu() - osr).abs().max().item() <= 1e-3


def test_boxes_to_locs_matches_the_script_arithmetic():
    """test_sr.py:118-134 restated with the script's own statements (Python floats, stored into an fp32 tensor)."""
    import torch
    from marconet_b200 import pipeline
    boxes = [[3.5, 2.0, 40.25, 30.0], [41.0, 1.0, 77.0, 31.0], [80.0, 0.0, 131.5, 33.0]]
    h, lq_width = 37, 512
    ref = torch.zeros(1, len(boxes) * 2).float()
    for i, box in enumerate(boxes):
        x1, y1, x2, y2 = box
        center = (x1 + x2) / 2.0
        width = (x2 - x1) / 2.0
        center_norm = center * 32.0 / h
        width_norm = width * 32.0 / h
        ref[0, 2 * i] = center_norm / lq_width
        ref[0, 2 * i + 1] = width_norm / lq_width
    assert torch.equal(pipeline.boxes_to_locs(boxes, h, lq_width), ref)


def test_round_half_even_is_cvround():
    from marconet_b200 import ops
    assert [ops.round_half_even(v) for v in (0.5, 1.5, 2.5, 2.4999, 186.45, 31.999999)] == [0, 2, 2, 2, 186, 32]
