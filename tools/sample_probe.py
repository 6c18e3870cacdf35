"""Times UpDownCaptioner.sample (batched diverse sampling) for growing samples-per-image on the bench dims."""
import sys
import time

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import sscvae  # noqa: E402


class _Vocab:
    def get_vocab_size(self, namespace="tokens"):
        return bench.DIMS["vocab_size"]

    def get_token_index(self, token, namespace="tokens"):
        return {"@@UNKNOWN@@": 0, "@@BOUNDARY@@": 1}[token]


def main():
    D = bench.DIMS
    dev = torch.device("cuda")
    torch.manual_seed(0)
    m = sscvae.UpDownCaptioner(_Vocab(), D["image_feature_size"], D["embedding_size"], D["hidden_size"],
                               D["attention_projection_size"], max_caption_length=20, beam_size=1, use_cbs=False,
                               z_space=D["z_space"], prior_std=1.0, latent_embedding="glove", sentiment_vae=1,
                               senti_prior_multip=0.5, cbs_simple=True, device=dev).to(dev)
    m.eval()
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    feats = torch.rand(n_img, 36, D["image_feature_size"], device=dev)
    sent = torch.zeros(n_img, 1, device=dev)
    for J in [int(a) for a in sys.argv[2:]] or [1, 4, 16, 100]:
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            out = m.sample(feats, sentiment=sent, n_samples=J)["predictions"]
            torch.cuda.synchronize()
            dt = time.time() - t0
            print(f"images {n_img} samples {J} call {it}: {dt * 1e3:.2f} ms -> {n_img * J / dt:.0f} captions/s, steps {out.shape[-1]}",
                  flush=True)


if __name__ == "__main__":
    main()
