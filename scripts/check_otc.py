"""Quick check of the tensor-core OccupancyMap path against the fp32 CUDA-core path (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from category_nerf_reconstruction_official_b200 import model, embedding, _lib
from category_nerf_reconstruction_official_b200._lib import check
dev = torch.device("cuda:0")
torch.manual_seed(0)
lib = _lib.load()
def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def occmap_bf16_emulation(params, emb, g_alpha, g_rgb):
    """fp32 torch restatement of the tensor-core OccupancyMap path with its bf16 roundings made explicit (operands of every GEMM
    are bf16: activations, deltas, weights, the biases that ride in the ones column; accumulation fp32).  params: the 14
    tensors in named_parameters() order.  Returns alpha, rgb, flat parameter gradient, g_emb."""
    Wi, bi, W1, b1, Wc, bc, W2, b2, wa, ba, Wl, bl, Wo, bo = [p.detach().float() for p in params]
    xg = bf(emb[:, :87]); xc = bf(emb[:, 87:])
    fc1 = bf(torch.relu(xg @ bf(Wi).T + bf(bi)))
    fc2 = bf(torch.relu(fc1 @ bf(W1).T + b1))
    fc3 = bf(torch.relu(fc2 @ bf(Wc[:, :128]).T + xg @ bf(Wc[:, 128:]).T + bf(bc)))
    pre4 = torch.relu(fc3 @ bf(W2).T + b2); fc4 = bf(pre4)
    alpha = 10.0 * (pre4 @ wa.T + ba).reshape(-1)
    preh = torch.relu(fc4 @ bf(Wl[:, :128]).T + xc @ bf(Wl[:, 128:]).T + bf(bl)); h = bf(preh)
    rgb = torch.sigmoid(preh @ Wo.T + bo)
    dp = g_rgb * rgb * (1 - rgb); draw = 10.0 * g_alpha
    hd = bf(torch.cat([dp, draw[:, None]], 1))
    dh = bf((dp @ Wo) * (h > 0))
    g_Wo = hd[:, :3].T @ h; g_bo = dp.sum(0)
    g_wa = (hd[:, 3:4].T @ fc4); g_ba = draw.sum().reshape(1)
    g_Wl = torch.cat([dh.T @ fc4, dh.T @ xc], 1); g_bl = dh.sum(0)
    d4 = bf((dh @ bf(Wl[:, :128]) + draw[:, None] * wa) * (fc4 > 0)); g_xc = dh @ bf(Wl[:, 128:])
    g_W2 = d4.T @ fc3; g_b2 = d4.sum(0)
    d3 = bf((d4 @ bf(W2)) * (fc3 > 0))
    g_Wc = torch.cat([d3.T @ fc2, d3.T @ xg], 1); g_bc = d3.sum(0)
    d2 = bf((d3 @ bf(Wc[:, :128])) * (fc2 > 0)); g_xg = d3 @ bf(Wc[:, 128:])
    g_W1 = d2.T @ fc1; g_b1 = d2.sum(0)
    d1 = bf((d2 @ bf(W1)) * (fc1 > 0))
    g_Wi = d1.T @ xg; g_bi = d1.sum(0)
    g_xg = g_xg + d1 @ bf(Wi)
    gp = torch.cat([t.reshape(-1) for t in (g_Wi, g_bi, g_W1, g_b1, g_Wc, g_bc, g_W2, g_b2, g_wa, g_ba, g_Wl, g_bl, g_Wo, g_bo)])
    return alpha, rgb, gp, torch.cat([g_xg, g_xc], 1)


def rel(a, b): return float((a - b).norm() / b.norm().clamp_min(1e-20))
for P in (128, 1000, 16800, 148 * 128 * 2 + 77):
    m = model.OccupancyMap(87, 42, hidden_size=128).to(dev)
    m.apply(model.init_weights)
    pe = embedding.UniDirsEmbed(max_deg=5, scale=5.0).to(dev)
    pts = (torch.rand(P, 3, device=dev) - 0.5) * 6
    with torch.no_grad():
        emb = pe(pts).contiguous()
    emb_r = emb.clone().requires_grad_(True)
    a_ref, c_ref = m(emb_r)
    ga = torch.randn(P, device=dev) * 0.1; gc = torch.randn(P, 3, device=dev)
    (a_ref.reshape(-1) * ga).sum().backward(retain_graph=True)
    (c_ref.reshape(-1, 3) * gc).sum().backward()
    gp_ref = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
    flat = model.flat_params(m).detach().contiguous()
    ws = torch.empty(lib.cnf_occmap_tc_workspace_bytes(P), dtype=torch.uint8, device=dev)
    alpha = torch.empty(P, device=dev); rgb = torch.empty(P, 3, device=dev)
    gp = torch.empty_like(flat); ge = torch.zeros(P, 129, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def fwd(): check(lib.cnf_occmap_tc_fwd(P, flat.data_ptr(), emb.data_ptr(), alpha.data_ptr(), rgb.data_ptr(), ws.data_ptr(), ws.numel(), st))
    def bwd(): check(lib.cnf_occmap_tc_bwd(P, flat.data_ptr(), rgb.data_ptr(), ga.data_ptr(), gc.data_ptr(), ws.data_ptr(), ws.numel(), gp.data_ptr(), ge.data_ptr(), st))
    fwd(); bwd()
    torch.cuda.synchronize()
    ea = float((alpha - a_ref.detach().reshape(-1)).abs().max()); ec = float((rgb - c_ref.detach().reshape(-1, 3)).abs().max())
    print("P=%d: alpha max abs err %.3e (range %.2f..%.2f)  rgb max abs err %.3e" % (P, ea, float(a_ref.min()), float(a_ref.max()), ec))
    names = [n for n, _ in m.named_parameters()]
    off = 0; worst = (0, "")
    for n, p in m.named_parameters():
        k = p.numel(); r = rel(gp[off:off + k], p.grad.reshape(-1)); off += k
        if r > worst[0]: worst = (r, n)
    print("   grads: whole rel L2 %.4f, worst tensor %s %.4f, g_emb rel L2 %.4f, finite %s" %
          (rel(gp, gp_ref), worst[1], worst[0], rel(ge, emb_r.grad), bool(torch.isfinite(gp).all())))
    a_e, c_e, gp_e, ge_e = occmap_bf16_emulation(list(m.parameters()), emb, ga, gc)
    print("   vs bf16 emulation: alpha max abs %.2e rgb max abs %.2e  grads whole rel L2 %.2e  g_emb rel L2 %.2e" %
          (float((alpha - a_e).abs().max()), float((rgb - c_e).abs().max()), rel(gp, gp_e), rel(ge, ge_e)))
    off = 0
    for n, p in m.named_parameters():
        k = p.numel(); print("      %-28s rel %.2e" % (n, rel(gp[off:off + k], gp_e[off:off + k]))); off += k
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print("   %s: %.1f us" % (name, min(ts) * 1e3))
