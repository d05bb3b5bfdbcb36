"""Device pre/post-processing against the oracle restatement of OpenCV / torchvision (oracle/image_ops.py, pinned
against the real cv2 in tests/test_oracle_image.py).  Byte / integer work: bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _img(h, w, seed, binary=False):
    rng = np.random.default_rng(seed)
    if binary:
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("h,w,seed", [(48, 300, 0), (20, 260, 1), (79, 460, 2), (32, 512, 3), (131, 97, 4), (9, 33, 5), (64, 1024, 6),
                                      (33, 17, 7), (4, 5, 8), (200, 3000, 9)])
def test_preprocess_bit_exact(h, w, seed):
    from marconet_b200 import ops
    from oracle import image_ops
    dev = torch.device("cuda:0")
    img = _img(h, w, seed, binary=(seed % 3 == 2))
    ref_lq, ref_w = image_ops.preprocess_lq(img)
    ref_small = image_ops.resize_cubic_u8(img, 32 / h, 32 / h)
    lq, lq_w, small = ops.preprocess_lq(torch.from_numpy(img).to(dev), return_resized=True)
    assert lq_w == ref_w and tuple(lq.shape) == (1, 3, 32, 512)
    assert np.array_equal(small.cpu().numpy(), ref_small), int((small.cpu().numpy() != ref_small).sum())
    assert np.array_equal(lq.cpu().numpy(), ref_lq)


def test_preprocess_matches_live_cv2_when_ipp_is_off():
    cv2 = pytest.importorskip("cv2")
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    was = cv2.ipp.useIPP()
    cv2.ipp.setUseIPP(False)
    try:
        for seed in range(6):
            rng = np.random.default_rng(100 + seed)
            h, w = int(rng.integers(10, 120)), int(rng.integers(30, 900))
            img = _img(h, w, 200 + seed, binary=bool(seed % 2))
            ref = cv2.resize(img, (0, 0), fx=32 / h, fy=32 / h, interpolation=cv2.INTER_CUBIC)
            if ref.shape[1] > 512:
                continue
            _, lq_w, small = ops.preprocess_lq(torch.from_numpy(img).to(dev), return_resized=True)
            assert lq_w == ref.shape[1] and np.array_equal(small.cpu().numpy(), ref)
    finally:
        cv2.ipp.setUseIPP(was)


def test_preprocess_rejects_wide_lines_and_cpu_tensors():
    from marconet_b200 import ops
    dev = torch.device("cuda:0")
    with pytest.raises(ValueError):
        ops.preprocess_lq(torch.from_numpy(_img(32, 600, 0)).to(dev))
    with pytest.raises(RuntimeError):
        ops.preprocess_lq(torch.from_numpy(_img(32, 100, 0)))


def test_postprocess_bit_exact_on_strided_views():
    from marconet_b200 import ops
    from oracle import image_ops
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4)
    sr = (torch.rand(2, 3, 128, 256, generator=g) * 2.6 - 1.3)
    sr[0, 0, 0, :6] = torch.tensor([-1.0, 1.0, 0.0, 1.0 / 255 - 1.0, 0.00196, -0.00196])
    ref = image_ops.postprocess_sr(sr.numpy())
    out = ops.postprocess_sr(sr.to(dev))
    assert np.array_equal(out.cpu().numpy(), ref)
    cl = sr.to(dev).contiguous(memory_format=torch.channels_last)          # what TSPSRNet returns: an NCHW view of NHWC memory
    assert np.array_equal(ops.postprocess_sr(cl).cpu().numpy(), ref)


def test_restore_image_end_to_end(gpu_models, checkpoints):
    """uint8 image in, uint8 image out: device pre/post-processing around the three modules equals the oracle's
    pre-processing -> oracle nets -> oracle post-processing within one grey level where the nets' 1e-3 tolerance crosses a
    rounding boundary."""
    from marconet_b200 import pipeline
    from oracle import image_ops, restate
    img = _img(40, 500, 11)
    labels = [5, 17, 300, 4242]
    boxes = [[20 + 110 * i, 4, 100 + 110 * i, 36] for i in range(4)]
    res = pipeline.restore_image(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"], img, labels, boxes)
    lq, lq_w = image_ops.preprocess_lq(img)
    assert res["lq_width"] == lq_w and np.array_equal(res["lq"].cpu().numpy(), lq)
    lq_t = torch.from_numpy(lq)
    locs = pipeline.boxes_to_locs(boxes, 40, 512)
    _, _, w = restate.encoder_forward(checkpoints["encoder"], lq_t)
    lab = torch.tensor(labels).reshape(-1, 1)
    _, f64, f32_ = restate.tspgan_forward(checkpoints["tspgan"], w[:1].repeat(4, 1), lab)
    sr = restate.tspsr_forward(checkpoints["sr"], lq_t, [f64], [f32_], locs)
    ref = image_ops.postprocess_sr(sr.numpy())[0, :, :res["sr_u8"].shape[1]]
    got = res["sr_u8"].cpu().numpy()
    assert got.shape == ref.shape == (128, 1600, 3)
    diff = np.abs(got.astype(int) - ref.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.15, (diff.max(), (diff != 0).mean())
    assert float((res["sr"].cpu() - sr).abs().max()) <= 1e-3


def test_restore_image_vs_reference_script_png(gpu_models):
    """Against the SR row of the PNG written by the UNMODIFIED reference test_sr.py on the CPU (tests/golden/script_sr_row.npz,
    made by oracle/make_golden_script.py; the oracle pipeline reproduces it byte for byte, tests/test_oracle_image.py): the device
    path may differ by one grey level where the nets' <= 1e-3 error crosses a rounding boundary."""
    import os
    from marconet_b200 import pipeline
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "script_sr_row.npz"))
    stride = int(g["stride"])
    res = pipeline.restore_image(gpu_models["encoder"], gpu_models["tspgan"], gpu_models["sr"], np.ascontiguousarray(g["image_rgb"]),
                                 g["labels"].tolist(), g["boxes"].tolist())
    got = res["sr_u8"].cpu().numpy()
    assert got.shape == (128, 1024, 3) and res["lq_width"] == 256
    diff = np.abs(got[::stride, ::stride].astype(int) - g["sr_row"].astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.15, (int(diff.max()), float((diff != 0).mean()))
