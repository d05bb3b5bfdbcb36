"""Module-level CUDA graphs behind the reference-facing forward() calls (marconet_b200/models/networks.py::_PackedModule):
the first call with a signature runs eagerly, the second records, later ones replay -- results must be bit-identical across all of
them, must never alias a later call's results, and errors must surface as the same exceptions."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fresh(checkpoints, dev):
    from marconet_b200.models import networks
    out = {}
    for key, cls in (("tspgan", networks.TSPGAN), ("encoder", networks.TextContextEncoderV2), ("sr", networks.TSPSRNet)):
        m = cls()
        m.load_state_dict(checkpoints[key], strict=True)
        out[key] = m.eval().to(dev)
    return out


def test_replays_are_bit_identical_to_the_eager_first_call_and_do_not_alias(checkpoints):
    from marconet_b200 import ops
    from oracle import synth
    assert ops.MODULE_GRAPHS
    dev = torch.device("cuda:0")
    nets = _fresh(checkpoints, dev)
    lqs = [synth.make_lq(1, s).to(dev) for s in (0, 1, 0, 0)]
    labels = synth.make_labels(4, 3)                      # CPU labels, like test_sr.py:180
    locs = synth.make_locs(1, 4, ragged=True, seed=2).to(dev)
    results = []
    for lq in lqs:                                        # call 1 eager, call 2 records + replays, calls 3-4 replay
        logits, locs_lr, w = nets["encoder"](lq)
        img, f64, f32_ = nets["tspgan"](styles=w.repeat(4, 1), labels=labels, noise=None)
        sr = nets["sr"](lq, [f64], [f32_], locs)
        results.append(dict(logits=logits, w=w, img=img, f64=f64, f32=f32_, sr=sr))
    assert len(nets["encoder"]._mg) == 1 and len(nets["tspgan"].TextGenerator._mg) == 1 and len(nets["sr"]._mg) == 1
    for k in results[0]:
        assert torch.equal(results[0][k], results[2][k]), f"replay differs from the eager call: {k}"     # same input (seed 0)
        assert torch.equal(results[0][k], results[3][k]), k
        assert results[2][k].data_ptr() != results[3][k].data_ptr(), f"results of two calls alias: {k}"
    assert not torch.equal(results[0]["sr"], results[1]["sr"])            # seed 1 really is another line
    # channels_last priors stay zero-copy inputs of the decoder
    assert results[3]["f64"].permute(0, 2, 3, 1).is_contiguous()


def test_graphed_modules_raise_like_the_eager_ones(checkpoints):
    from oracle import synth
    dev = torch.device("cuda:0")
    nets = _fresh(checkpoints, dev)
    styles = synth.make_styles(3, 1).to(dev)
    good = torch.tensor([[5], [17], [6000]])
    for _ in range(3):
        img, f64, f32_ = nets["tspgan"](styles=styles, labels=good.to(dev), noise=None)       # device labels: flag read back
    with pytest.raises(IndexError):
        nets["tspgan"](styles=styles, labels=torch.tensor([[5], [-1], [3]]).to(dev), noise=None)   # replayed graph, device-side check
    with pytest.raises(IndexError):
        nets["tspgan"](styles=styles, labels=torch.tensor([[5], [6736], [3]]), noise=None)         # CPU labels: host check
    again = nets["tspgan"](styles=styles, labels=good.to(dev), noise=None)
    assert torch.equal(again[0], img)
    lq = synth.make_lq(1, 4).to(dev)
    locs = synth.make_locs(1, 3).to(dev)
    for _ in range(3):
        sr = nets["sr"](lq, [f64], [f32_], locs)
    bad = locs.clone()
    bad[0, 2] = -0.2                                      # empty window: the reference fails on the empty slice (networks.py:443)
    with pytest.raises(RuntimeError):
        nets["sr"](lq, [f64], [f32_], bad)
    assert torch.equal(nets["sr"](lq, [f64], [f32_], locs), sr)


def test_signatures_are_cached_separately_and_weights_changes_invalidate(checkpoints):
    from oracle import synth
    dev = torch.device("cuda:0")
    nets = _fresh(checkpoints, dev)
    gen = nets["tspgan"]
    outs = {}
    for rep in range(3):
        for n in (2, 5):
            img, _, _ = gen(styles=synth.make_styles(n, n).to(dev), labels=synth.make_labels(n, n), noise=None)
            if rep == 0:
                outs[n] = img
            else:
                assert torch.equal(outs[n], img)
    assert len(gen.TextGenerator._mg) == 2
    # an in-place parameter update re-packs the weights and drops the recorded graphs
    with torch.no_grad():
        gen.TextGenerator.conv1.bias.add_(0.5)
    img2, _, _ = gen(styles=synth.make_styles(2, 2).to(dev), labels=synth.make_labels(2, 2), noise=None)
    assert not torch.equal(img2, outs[2]) and not gen.TextGenerator._mg
