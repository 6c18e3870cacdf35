"""Per-kernel-class time of one CBS decode call, one greedy call and one batched sampling call (library
instrumentation, no CUDA graphs)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import sscvae
from sscvae import _lib
import bench

dev = torch.device("cuda", 0)
torch.manual_seed(0)

class V:
    def get_vocab_size(self, namespace="tokens"): return bench.DIMS["vocab_size"]
    def get_token_index(self, token, namespace="tokens"): return {"@@UNKNOWN@@": 0, "@@BOUNDARY@@": 1}[token]

D = bench.DIMS
def build(beam, cbs):
    m = sscvae.UpDownCaptioner(V(), D["image_feature_size"], D["embedding_size"], D["hidden_size"], D["attention_projection_size"],
                               max_caption_length=20, beam_size=beam, use_cbs=cbs, min_constraints_to_satisfy=2, z_space=D["z_space"],
                               prior_std=1.0, simple_vae=False, latent_embedding="glove", sentiment_vae=1, senti_prior_multip=0.5,
                               cbs_simple=True, device=dev).to(dev)
    m.eval()
    return m

g = torch.Generator().manual_seed(7)
for name, beam, cbs, n_img in (("cbs_beam5", 5, True, 64), ("greedy", 1, False, 256)):
    feats = torch.rand(n_img, 36, 2048, generator=g).to(dev)
    sent = torch.randint(-1, 2, (n_img, 1), generator=g).float().to(dev)
    m = build(beam, cbs)
    kw = {}
    if cbs:
        kw = dict(fsm=bench.synthetic_fsm(n_img, D["vocab_size"], [[11, 12], [57], [300, 301, 302]]).to(dev),
                  num_constraints=torch.full((n_img,), 3, dtype=torch.long, device=dev))
    for _ in range(2):
        m(feats, None, None, sentiment=sent, **kw)
    _lib.profile(True)
    m(feats, None, None, sentiment=sent, **kw)
    rep = _lib.profile_report()
    _lib.profile(False)
    tot = sum(v["ms"] for v in rep.values())
    print(name, "images", n_img, "total ms", round(tot, 3))
    for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
        print(f"  {k:20s} {v['count']:5d} {v['ms']:8.3f} ms  {v['ms'] / tot * 100:5.1f}%")

# batched diverse sampling: 64 images x 100 samples in one call
n_img, J = 64, 100
feats = torch.rand(n_img, 36, 2048, generator=g).to(dev)
sent = torch.randint(-1, 2, (n_img, 1), generator=g).float().to(dev)
m = build(1, False)
for _ in range(2):
    m.sample(feats, sentiment=sent, n_samples=J)
_lib.profile(True)
m.sample(feats, sentiment=sent, n_samples=J)
rep = _lib.profile_report()
_lib.profile(False)
tot = sum(v["ms"] for v in rep.values())
print("sampling_batched images", n_img, "samples", J, "total ms", round(tot, 3))
for k, v in sorted(rep.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"  {k:20s} {v['count']:5d} {v['ms']:8.3f} ms  {v['ms'] / tot * 100:5.1f}%")
