"""CUDA-graph capture of the line pipeline and the device-side checks it relies on.

* the window integers computed by mn_char_windows are bit-identical to the host restatement (which is pinned against the
  reference's golden windows in test_gpu_models.py::test_window_integers_bit_exact),
* the modules produce bit-identical outputs inside ops.deferred_checks (no host round trip) and outside it,
* a replayed graph reproduces the eager module calls bit for bit, for new inputs copied into its static buffers,
* bad labels / empty windows surface as the same exceptions, after the replay.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _flag(dev):
    return torch.zeros(1, dtype=torch.int32, device=dev)


@pytest.mark.parametrize("width,half", [(512, 16), (1024, 32)])
def test_device_windows_bit_exact(width, half):
    from marconet_b200 import ops
    from marconet_b200.models.networks import char_windows
    from oracle.make_golden import case_inputs
    dev = torch.device("cuda:0")
    cases = []
    inp = case_inputs("ragged")
    cases.append((inp["locs"], [l.shape[0] for l in inp["labels"]]))
    g = torch.Generator().manual_seed(5)
    # random centres including the clipped ends, products that land next to an integer, and overlapping windows
    locs = torch.rand(7, 32, generator=g)
    locs[0, 0] = 0.0
    locs[0, 2] = 1.0
    locs[1, 0::2] = (torch.arange(16, dtype=torch.float32) * 31 + 3) / width
    locs[2, 0::2] = torch.nextafter(torch.arange(1, 17, dtype=torch.float32) * 29 / width, torch.tensor(0.0))
    cases.append((locs, [16, 16, 16, 3, 0, 9, 1]))
    for locs_h, counts in cases:
        wins, valid, owner = char_windows(locs_h, counts, width, half)
        first = torch.tensor(np.concatenate([[0], np.cumsum(counts)]), dtype=torch.int32).to(dev)
        flag = _flag(dev)
        win_d, valid_d, owner_d = ops.char_windows(locs_h.to(dev), first, counts, width, half, flag)
        assert int(flag.item()) == 0
        assert np.array_equal(win_d.cpu().numpy(), np.asarray(wins, dtype=np.int32).reshape(-1, 4))
        assert np.array_equal(valid_d.cpu().numpy(), np.asarray(valid, dtype=np.int32))
        assert np.array_equal(owner_d.cpu().numpy(), np.asarray(owner, dtype=np.int32))


def test_device_windows_flag_empty_window():
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    locs = torch.full((1, 4), 0.5)
    locs[0, 2] = 1.5                        # centre 768 of 512: x1 = 752 >= W -> the reference dies on the empty slice
    first = torch.tensor([0, 2], dtype=torch.int32).to(dev)
    flag = _flag(dev)
    win, valid, owner = ops.char_windows(locs.to(dev), first, [2], 512, 16, flag)
    assert int(flag.item()) == ops.ERR_WINDOW
    assert valid.cpu().tolist() == [32, 0]
    with pytest.raises(RuntimeError):
        ops.raise_deferred(int(flag.item()))


def test_check_labels_clamps_and_flags():
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    flag = _flag(dev)
    lab = torch.tensor([0, 6735, 17], dtype=torch.int64, device=dev)
    out = ops.check_labels(lab, 6736, flag)
    assert int(flag.item()) == 0 and torch.equal(out, lab)
    bad = torch.tensor([-3, 6736, 17], dtype=torch.int64, device=dev)
    out = ops.check_labels(bad, 6736, flag)
    assert int(flag.item()) == ops.ERR_LABEL and out.cpu().tolist() == [0, 6735, 17]
    with pytest.raises(IndexError):
        ops.raise_deferred(int(flag.item()))


def test_deferred_modules_match_eager_bitwise(gpu_models):
    from marconet_b200 import ops
    from oracle.make_golden import case_inputs
    dev = torch.device("cuda:0")
    inp = case_inputs("ragged")
    lq = inp["lq"].to(dev)
    labels = torch.cat(inp["labels"]).to(dev)
    counts = [l.shape[0] for l in inp["labels"]]
    locs = inp["locs"].to(dev)
    styles = torch.randn(labels.shape[0], 512, generator=torch.Generator().manual_seed(3)).to(dev)

    def run():
        img, f64, f32_ = gpu_models["tspgan"](styles=styles, labels=labels, noise=None)
        o, p64, p32 = 0, [], []
        for n in counts:
            p64.append(f64[o:o + n]); p32.append(f32_[o:o + n]); o += n
        return img, gpu_models["sr"](lq, p64, p32, locs)

    img_e, sr_e = run()
    flag = _flag(dev)
    with ops.deferred_checks(flag):
        img_d, sr_d = run()
    assert int(flag.item()) == 0
    assert torch.equal(img_e, img_d) and torch.equal(sr_e, sr_d)


def test_graph_replay_matches_eager(gpu_models):
    from marconet_b200.graph import GraphedLines
    from marconet_b200.testing import synth
    dev = torch.device("cuda:0")
    lines, chars = 2, 4
    g = GraphedLines(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"], lines=lines, chars=chars)
    assert g.launches > 50
    for seed in (11, 12):
        lq = synth.make_lq(lines, seed)
        labels = torch.cat([synth.make_labels(chars, seed + b) for b in range(lines)])
        locs = synth.make_locs(lines, chars, ragged=(seed == 12), seed=seed)
        sr = g(lq, labels, locs)                 # host tensors -> static buffers -> replay
        g.check()
        lqd = lq.to(dev)
        _, _, w = gpu_models["encoder"](lqd)
        _, f64, f32_ = gpu_models["tspgan"](styles=w.repeat_interleave(chars, dim=0), labels=labels, noise=None)
        ref = gpu_models["sr"](lqd, [f64[b * chars:(b + 1) * chars] for b in range(lines)],
                               [f32_[b * chars:(b + 1) * chars] for b in range(lines)], locs.to(dev))
        assert torch.equal(sr, ref), float((sr - ref).abs().max())
        assert torch.equal(g.outputs["w"], w)


def test_graph_deferred_errors(gpu_models):
    from marconet_b200.graph import GraphedLines
    from marconet_b200.testing import synth
    g = GraphedLines(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"], lines=1, chars=2)
    lq, locs = synth.make_lq(1, 1), synth.make_locs(1, 2)
    g(lq, torch.tensor([[3], [7000]]), locs)
    with pytest.raises(IndexError):
        g.check()
    bad = locs.clone()
    bad[0, 0] = 1.5
    g(lq, torch.tensor([[3], [4]]), bad)
    with pytest.raises(RuntimeError):
        g.check()
    g(lq, torch.tensor([[3], [4]]), locs)      # the flag is cleared inside the graph: a good call after a bad one is clean
    g.check()
    assert torch.isfinite(g.outputs["sr"]).all()
