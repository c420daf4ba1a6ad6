"""Diagnostic (GPU): error of each precision mode on BASELINE config 1 (2DRB_1) against the CPU oracle -- the numbers
quoted in tests/test_gpu_model.py's header.  Lives under tests/ because it uses the oracle as the checker."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import rnamsm_b200 as pkg
from oracle import msa_ref as O
def argmax_agreement(a, b):
    a = torch.as_tensor(a).reshape(-1, a.shape[-1]); b = torch.as_tensor(b).reshape(-1, b.shape[-1])
    return float((a.argmax(-1) == b.argmax(-1)).float().mean())
vocab = pkg.Vocab(pkg.Alphabet())
g = np.load(os.path.join(ROOT, "tests", "golden", "2DRB_1.npz"))
tokens = torch.from_numpy(g["tokens"].astype(np.int64)).cuda()
sd = O.make_weights(int(g["wseed"]), sharpen=float(g["sharpen"]))
for prec in ("bf16_pure", "bf16", "fp16", "fp32"):
    m = pkg.MSATransformer(vocab, num_layers=10, precision=prec); m.load_state_dict(sd, strict=True); m = m.eval().cuda()
    out = m(tokens, repr_layers=[10], need_head_weights=True, want_logits=False)
    emb, atp = pkg.extract_features(out, vocab, 10)
    print("2DRB", prec, "emb %.3e atp %.3e argmax %.4f" % (O.rel_err(emb, g["emb"]), O.rel_err(atp, g["atp"]), argmax_agreement(atp, g["atp"])), flush=True)
sd = O.make_weights(42, num_layers=3, sharpen=3.0)
tokens = O.make_tokens(96, 200, 8)
ref = O.forward(sd, tokens, repr_layers=[3], need_head_weights=True, num_layers=3, want_logits=False)
for prec in ("bf16_pure", "bf16", "fp16", "fp32"):
    m = pkg.MSATransformer(vocab, num_layers=3, precision=prec); m.load_state_dict(sd, strict=True); m = m.eval().cuda()
    out = m(tokens.cuda(), repr_layers=[3], need_head_weights=True, want_logits=False)
    print("mid", prec, "rep %.3e att %.3e argmax %.4f" % (O.rel_err(out["representations"][3].cpu(), ref["representations"][3]), O.rel_err(out["row_attentions"].cpu(), ref["row_attentions"]), argmax_agreement(out["row_attentions"].cpu(), ref["row_attentions"])), flush=True)
