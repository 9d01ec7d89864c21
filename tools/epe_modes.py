import sys, types, numpy as np, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests/golden')
import anystereo_b200 as A
from oracle import hotpath_oracle as O
for fam in ("igev", "raft"):
    g = dict(np.load('/root/repo/tests/golden/model_%s_boundary.npz' % fam))
    net = [torch.from_numpy(g["net%d" % i]).cuda() for i in range(3)]
    inp = [[torch.from_numpy(g["inp%d_%d" % (i, j)]).cuda() for j in range(3)] for i in range(3)]
    f1, f2 = torch.from_numpy(g["f1"]).cuda(), torch.from_numpy(g["f2"]).cuda()
    args = types.SimpleNamespace(corr_levels=2 if fam == "igev" else 4, corr_radius=4, n_gru_layers=3)
    cls = A.BasicMultiUpdateBlock if fam == "igev" else A.BasicMultiUpdateBlockRAFT
    m = cls(args, hidden_dims=[128] * 3)
    m.load_state_dict(O.make_update_block_params(162 if fam == "igev" else 36, seed=78 if fam == "igev" else 77), strict=True)
    m = m.cuda().eval()
    for eng in ("fp32", "bf16x3", "bf16", "fp16"):
        A.set_update_engine(eng); A.set_corr_mode("bf16x3" if eng == "fp16" else eng)
        if fam == "igev":
            d, _ = A.igev_iterations(m, f1, f2, torch.from_numpy(g["geo"]).cuda(), [t.clone() for t in net], inp, torch.from_numpy(g["init_disp"]).cuda(), int(g["iters"]))
        else:
            d, _ = A.raft_iterations(m, f1, f2, [t.clone() for t in net], inp, int(g["iters"]))
        e = (d.cpu() - torch.from_numpy(g["disp_lowres"])).abs() * 4
        print(fam, eng, "EPE mean %.2e max %.2e px" % (float(e.mean()), float(e.max())))
