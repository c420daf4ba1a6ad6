"""torchrun debug driver for the fused sharded schedule: prints progress, dumps stacks if it stalls."""
import faulthandler, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
faulthandler.dump_traceback_later(50, exit=True)
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
import rnamsm_b200 as pkg
from rnamsm_b200 import sharded
from oracle import msa_ref as O
def log(*a):
    print(f"[r{rank} {time.time() % 1000:.2f}]", *a, flush=True)
vocab = pkg.Vocab(pkg.Alphabet())
m = pkg.MSATransformer(vocab, num_layers=2, precision="fp16")
m.load_state_dict(O.make_weights(9, num_layers=2, sharpen=2.0), strict=True)
m = m.eval().cuda()
tokens = O.make_tokens(64, 96, 4).cuda()
log("model ready")
ref = m(tokens, repr_layers=[2], need_head_weights=True, want_logits=False)
torch.cuda.synchronize(); log("reference forward done")
pb = sharded.PeerBuffer(1 << 20)
log("peer buffer ok", [hex(p) for p in pb.ptrs])
t = pb.tensor(torch.float32, (16,)); t.fill_(rank + 1.0); torch.cuda.synchronize()
if world > 1:
    dist.barrier()
    import ctypes as C
    # read the peer's buffer through the mapped pointer with a cudaMemcpy via torch: wrap peer ptr
    class H: pass
    h = H(); h.__cuda_array_interface__ = {"shape": (64,), "typestr": "|u1", "data": (pb.ptrs[(rank + 1) % world], False), "version": 2}
    peer = torch.as_tensor(h, device="cuda").view(torch.float32)
    log("peer value", float(peer[0].item()))
out = sharded.sharded_forward(m, tokens, fused=True)
torch.cuda.synchronize(); log("fused forward done")
r0, r1 = out["row_shard"]
log("rep err", O.rel_err(out["representations"][2].cpu(), ref["representations"][2][:, r0:r1].cpu()))
if rank == 0:
    log("map err", O.rel_err(out["row_attentions"].cpu(), ref["row_attentions"].cpu()))
if world > 1:
    dist.destroy_process_group()
log("done")
