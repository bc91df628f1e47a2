import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo" if os.path.exists("/root/repo/tests") else ".")
import tracknetv3_b200 as T
from tests import gpu_util as G
g = np.load("tests/golden/tracknet_step.npz", allow_pickle=False)
torch.manual_seed(int(g["seed"]))
m = T.TrackNet(27, 8).to("cuda").train()
x, y = torch.from_numpy(g["x"]).cuda(), torch.from_numpy(g["y"]).cuda()
yp = m(x); loss = T.WBCELoss(yp, y); loss.backward()
print("heatmap max-abs", np.abs(yp.detach().cpu().numpy() - g["y_pred"]).max(), "x", tuple(x.shape))
ours = dict(m.named_parameters())["up_block_3.conv_2.conv.weight"].grad.cpu().double()
ref = torch.from_numpy(g["grad_last"]).double()
e = (ours - ref).abs() / ref.abs().max()
print("last conv weight grad: max rel err", e.max().item(), "elements > 1e-3:", (e > 1e-3).sum().item(), "of", e.numel(), " > 1e-4:", (e > 1e-4).sum().item(), "median", e.median().item())
idx = torch.topk(e.flatten(), 8).indices
for i in idx:
    co, r = divmod(i.item(), 64 * 9); ci, t = divmod(r, 9)
    print(f"  co {co} ci {ci} tap {t}: ours {ours.flatten()[i].item():+.6e} ref {ref.flatten()[i].item():+.6e} err {e.flatten()[i].item():.3e}")
