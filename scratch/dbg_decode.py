import sys, numpy as np, torch
sys.path.insert(0, "/root/repo/computer-vision-models_b200"); sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import synth
from oracle import decode_np
from oracle.layout import make_layout
from cvmhot.models.centernet import CenternetParams
from cvmhot.models.centernet.post_processing import decode_topk
H, W, C, B, K = 128, 384, 10, 4, 100
Lo = make_layout(H, W, C, "N")
data = synth.make_batch(Lo, 4, B)
yp = data["y_pred"]
ref = decode_np.decode_topk(Lo, yp, K)
p = CenternetParams(C, True); p.INPUT_HEIGHT, p.INPUT_WIDTH = H*2, W*2
for rep in range(3):
    out = decode_topk(torch.from_numpy(yp).cuda(), p, K=K)
    got = out["flat"].cpu().numpy()
    for b in range(B):
        miss = sorted(set(ref["flat"][b].tolist()) - set(got[b].tolist()))
        extra = sorted(set(got[b].tolist()) - set(ref["flat"][b].tolist()))
        if miss or extra:
            def d(f): return (f // C // W, (f // C) % W, f % C, ((f // C) // 512, (f // C) % 512))
            print(rep, b, "n_missing", len(miss), [(d(f), round(float(yp[b].reshape(-1, Lo.Cp)[f // C, f % C]), 3)) for f in miss][:6], "extra", [d(f) for f in extra][:3])
print("done")
