import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sella_b200.internal import BatchedInternals
G = np.load(os.path.join(ROOT, "tests/golden/rotation.npz"))
dev = torch.device("cuda:0")
for i in range(int(G["ncases"])):
    ref, pos = G["ref%d" % i], G["pos%d" % i]
    ints = BatchedInternals(len(ref), rotation_ref=ref)
    x = torch.from_numpy(pos.ravel()[None].copy()).to(dev)
    q, B = ints.calc(x, jacobian=True)
    print(i, len(ref), "q dev", ints.qprev.cpu().numpy()[0].round(6), "q ref", G["q%d" % i].round(6), "val err", np.abs(q.cpu().numpy()[0] - G["val%d" % i]).max(),
          "jac err", np.abs(B.cpu().numpy()[0] - G["jac%d" % i]).max())
