"""N4 (RND / NovelD): novelty and predictor update, fp32 FMA path against the bf16 tensor path, at the reference's
update-batch sizes and at a large batch."""
import sys, torch
sys.path.insert(0, '.')
from oracle import port
from ddiffpg_b200 import FusedRNDTrainer, RNDModel

def timed(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(n): fn()
    ev[1].record(); torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / n

p = port.init_rnd_params(7)
for B in (4096, 8192, 131072):
    x = torch.randn(B, 69, device="cuda")
    row = [f"rows {B:7d}"]
    for prec in ("fp32", "bf16"):
        m = RNDModel(69, precision=prec); m.load_state_dict(p); m = m.to("cuda")
        nov, upd = timed(lambda: m.novelty(x)), timed(lambda: m.loss_and_grads(x))
        # the whole IntrinsicM.update: + clip + AdamW, torch's own tail against the fused trainer (CUDA graph)
        opt = torch.optim.AdamW(m.predictor.parameters(), 1e-4)
        def torch_step():
            loss, g = m.loss_and_grads(x)
            off = 0
            for q in m.predictor.parameters():
                q.grad = g[off:off + q.numel()].view(q.shape); off += q.numel()
            torch.nn.utils.clip_grad_norm_(m.predictor.parameters(), 1.0); opt.step()
        t_torch = timed(torch_step)
        tr = FusedRNDTrainer(m, graph=True)
        t_fused = timed(lambda: tr.step(x))
        tr.close()
        row.append(f"{prec}: novelty {nov:.3f} ms, loss+grads {upd:.3f} ms, update with torch tail {t_torch:.3f} ms, fused trainer {t_fused:.3f} ms")
    print("   ".join(row))
