"""GPU parity: the S3Gaussian EMD deformation network (K1d) vs the oracle (itself pinned to the reference module)."""
import os

import numpy as np
import pytest
import torch

from tests.util import rel_err, rel_l2

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("M,K,Nout,ri,ro", [(1000, 64, 64, True, True), (777, 132, 64, False, False),
                                             (513, 4, 64, False, False), (300, 64, 48, False, False),
                                             (129, 64, 1, False, False), (5000, 64, 3, True, False),
                                             (128, 64, 64, False, True), (1, 64, 64, True, True),
                                             # the bulk-staged weight gradient: ~32 chunks per CTA (ring phases), no tail
                                             # rows, one whole chunk + tail, Nout not a multiple of 4.  No output ReLU at
                                             # these sizes: among 2e7 outputs some lie within rounding of 0, and a mask
                                             # that flips there changes a whole gradient row (the mask itself is
                                             # per element and covered by the small cases)
                                             (300017, 64, 64, True, False), (65536, 4, 64, False, False),
                                             (40, 64, 3, False, False), (100000, 64, 48, True, False)])
def test_linear_layer(M, K, Nout, ri, ro):
    from emd_b200.mlp_ops import linear
    g = torch.Generator().manual_seed(M + K + Nout)
    X = torch.randn(M, K, generator=g).requires_grad_(True)
    W = (torch.randn(Nout, K, generator=g) / K ** 0.5).requires_grad_(True)
    b = torch.randn(Nout, generator=g).requires_grad_(True)
    xin = torch.relu(X) if ri else X
    Y = torch.nn.functional.linear(xin, W, b)
    Y = torch.relu(Y) if ro else Y
    v = torch.randn(M, Nout, generator=g)
    (Y * v).sum().backward()
    Xg, Wg, bg = [t.detach().cuda().requires_grad_(True) for t in (X, W, b)]
    Yg = linear(Xg, Wg, bg, relu_in=ri, relu_out=ro)
    assert float((Yg.detach().cpu() - Y.detach()).abs().max()) <= 2e-5 * max(1.0, float(Y.detach().abs().max()))
    (Yg * v.cuda()).sum().backward()
    for a, r, name in ((Xg.grad, X.grad, "dX"), (Wg.grad, W.grad, "dW"), (bg.grad, b.grad, "db")):
        assert rel_err(a, r) <= 2e-4, (name, rel_err(a, r))


def test_deformation_network_golden_and_grads():
    from emd_b200.emd_s3g import S3GDeformation
    from oracle import emd_s3g as S
    z = np.load(f"{G}/emd_s3g.npz")
    t = lambda a: torch.from_numpy(np.asarray(a))  # noqa: E731
    names = [k[len("w.deformation_net."):] for k in z.files if k.startswith("w.deformation_net.")]
    used = [n for n in names if not any(s in n for s in ("scales_deform", "rotations_deform"))]
    for ci, (tm, it, cam) in enumerate(z["cases"].tolist()):
        w_c = {n: t(z["w.deformation_net." + n]).clone().requires_grad_(True) for n in used}
        inp_c = {k: t(z[k]).clone().requires_grad_(True) for k in ("point", "opacity", "shs", "embeddings")}
        hex_c = t(z[f"c{ci}_hex"]).clone().requires_grad_(True)
        means, opac, shs, dd = S.deform(w_c, inp_c["point"], inp_c["opacity"], inp_c["shs"], inp_c["embeddings"],
                                        hex_c, float(np.float32(tm)), int(it), int(cam))
        dev = "cuda"
        w_g = {n: v.detach().to(dev).requires_grad_(True) for n, v in w_c.items()}
        inp_g = {k: v.detach().to(dev).requires_grad_(True) for k, v in inp_c.items()}
        hex_g = hex_c.detach().to(dev).requires_grad_(True)
        net = S3GDeformation(w_g)
        gm, gs, gr, go, gsh, gdd = net(inp_g["point"], t(z["scales"]).to(dev), t(z["rotations"]).to(dev),
                                       inp_g["opacity"], inp_g["shs"], float(np.float32(tm)), inp_g["embeddings"],
                                       int(it), int(cam), hex_g)
        # forward: against the oracle AND directly against the reference module's outputs
        for a, b_, name in ((gm, t(z[f"c{ci}_means"]), "means"), (go, t(z[f"c{ci}_opacity"]), "opacity"),
                            (gsh, t(z[f"c{ci}_shs"]), "shs")):
            assert float((a.detach().cpu() - b_).abs().max()) <= 5e-6, (ci, name)
        g = torch.Generator().manual_seed(ci)
        cot = {"m": torch.randn(means.shape, generator=g), "o": torch.randn(opac.shape, generator=g),
               "s": torch.randn(shs.shape, generator=g)}
        loss_c = (means * cot["m"]).sum() + (opac * cot["o"]).sum() + (shs * cot["s"]).sum()
        loss_g = (gm * cot["m"].to(dev)).sum() + (go * cot["o"].to(dev)).sum() + (gsh * cot["s"].to(dev)).sum()
        for br in ("coarse", "fine"):
            for key in ("dx", "do", "dshs", "feat"):
                assert float((gdd[br][key].detach().cpu() - t(z[f"c{ci}_{br}_{key}"])).abs().max()) <= 5e-6, (ci, br, key)
                c = torch.randn(dd[br][key].shape, generator=g)
                loss_c = loss_c + (dd[br][key].abs() * c).sum()      # the trainer's L1 regularisers touch these
                loss_g = loss_g + (gdd[br][key].abs() * c.to(dev)).sum()
        loss_c.backward()
        loss_g.backward()
        for n in used:
            gr_, gg_ = w_c[n].grad, w_g[n].grad
            if gr_ is None or float(gr_.abs().max()) == 0.0:
                continue
            assert gg_ is not None, n
            assert rel_err(gg_, gr_) <= 1e-3 and rel_l2(gg_, gr_) <= 1e-3, (ci, n, rel_err(gg_, gr_))
        for k in inp_c:
            assert rel_err(inp_g[k].grad, inp_c[k].grad) <= 1e-3, (ci, k)
        assert rel_err(hex_g.grad, hex_c.grad) <= 1e-3, (ci, "hex")
