"""Ad-hoc parity check of the whole path on a different synthetic checkpoint / input seed than the committed fixtures
(a script, not collected by pytest: python tests/parity_seed_check.py SEED NCHARS).  Lives under tests/ because it runs the oracle."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200.models import networks
from oracle import restate, synth

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
nchars = int(sys.argv[2]) if len(sys.argv) > 2 else 6
dev = torch.device("cuda:0")
sds = synth.make_checkpoints(seed)
nets = {}
for key, cls in (("tspgan", networks.TSPGAN), ("encoder", networks.TextContextEncoderV2), ("sr", networks.TSPSRNet)):
    m = cls(); m.load_state_dict(sds[key], strict=True); nets[key] = m.eval().to(dev)
lq = synth.make_lq(1, 100 + seed)
labels, locs = synth.make_labels(nchars, 50 + seed), synth.make_locs(1, nchars, ragged=True, seed=seed)
with torch.no_grad():
    logits, _, w = nets["encoder"](lq.to(dev))
    img, f64, f32_ = nets["tspgan"](styles=w.repeat(nchars, 1), labels=labels, noise=None)
    sr = nets["sr"](lq.to(dev), [f64], [f32_], locs.to(dev))
ol, _, ow = restate.encoder_forward(sds["encoder"], lq)
oi, o64, o32 = restate.tspgan_forward(sds["tspgan"], ow.repeat(nchars, 1), labels)
osr = restate.tspsr_forward(sds["sr"], lq, [o64], [o32], locs)
errs = dict(logits=(logits.cpu() - ol).abs().max().item(), w=(w.cpu() - ow).abs().max().item(), image=(img.cpu() - oi).abs().max().item(),
            fea64=(f64.cpu() - o64).abs().max().item(), fea32=(f32_.cpu() - o32).abs().max().item(), sr=(sr.cpu() - osr).abs().max().item())
print("seed", seed, "chars", nchars, {k: f"{v:.2e}" for k, v in errs.items()}, "argmax equal:", torch.equal(logits.argmax(-1).cpu(), ol.argmax(-1)))
