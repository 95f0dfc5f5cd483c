"""PyTorch-CPU (MKL / oneDNN) fp64 model of the scoring path -- the closest stand-in available here for Torch7 + MKL's
BLAS-backed nn.Linear / nn.FastLSTM (SURVEY.md 8d "PyTorch-CPU proxy").

TEST / BASELINE INFRASTRUCTURE ONLY (like everything under oracle/): bench.py's cpu_baseline leg times it beside the
C restatement; it is never imported from kprn_b200/.  Graph: SP/model/OneModel.lua:203-294 -- three LookupTable gathers +
JoinTable (FeatureEmbedding.lua:112-121), FastLSTM (gate blocks i,g,f,o; no masking), Linear(H,L) on the last step,
LogSumExp per (user,item) segment (module/LogSumExp.lua:13-30), Sigmoid, Select(classId)."""
import numpy as np
import torch


def forward(o, params, ids, seg, class_id=1):
    """o: Oracle (shapes/offsets); params flat fp64 numpy; ids [N,T,3] 1-based; seg CSR. Returns (path_logits, scores) numpy."""
    c = o.cfg
    p = torch.from_numpy(np.ascontiguousarray(params, dtype=np.float64))
    P = {k: p[o.offsets[k]:o.offsets[k] + int(np.prod(s))].reshape(s) for k, s in o.shapes().items()}
    idx = torch.from_numpy(np.asarray(ids, dtype=np.int64)) - 1
    N, H = idx.shape[0], c.H
    with torch.no_grad():
        # x W_x^T for all steps in one GEMM (what a BLAS-backed nn.Sequencer amortises to), then the recurrence
        x = torch.cat([P["E_t"][idx[:, :, 0]], P["E_e"][idx[:, :, 1]], P["E_r"][idx[:, :, 2]]], dim=2)     # [N,T,D]
        xw = torch.addmm(P["bx"], x.reshape(N * c.T, -1), P["Wx"].T).reshape(N, c.T, 4 * H)
        h = torch.zeros(N, H, dtype=torch.float64)
        cc = torch.zeros(N, H, dtype=torch.float64)
        WhT = P["Wh"].T.contiguous()
        for t in range(c.T):
            G = xw[:, t] + h @ WhT
            i, g = torch.sigmoid(G[:, :H]), torch.tanh(G[:, H:2 * H])
            f, og = torch.sigmoid(G[:, 2 * H:3 * H]), torch.sigmoid(G[:, 3 * H:])
            cc = i * g + f * cc
            h = og * torch.tanh(cc)
        s = torch.addmm(P["bo"], h, P["Wo"].T)
        sc = s[:, class_id - 1]
        segt = torch.from_numpy(np.asarray(seg, dtype=np.int64))
        S = segt.numel() - 1
        sid = torch.repeat_interleave(torch.arange(S), segt[1:] - segt[:-1])
        m = torch.full((S,), -float("inf"), dtype=torch.float64).scatter_reduce(0, sid, sc, reduce="amax")
        z = torch.zeros(S, dtype=torch.float64).index_add_(0, sid, torch.exp(sc - m[sid]))
        lse = m + torch.log(z)
        scores = torch.sigmoid(lse)
    return s.numpy(), scores.numpy()
