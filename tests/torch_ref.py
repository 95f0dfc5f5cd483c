"""Independent fp64 PyTorch autograd model of the path (second opinion on the C oracle).

FastLSTM gate order is [i, g, f, o] (Element-Research rnn); written here with explicit
matmuls rather than nn.LSTM so the autograd graph is transparent.
"""
import torch


def forward(o, params, ids, seg, class_id=1):
    """o: oracle.Oracle (for shapes/offsets); params: flat fp64 torch tensor (requires_grad ok)."""
    c = o.cfg
    P = {k: params[o.offsets[k]:o.offsets[k] + int(torch.tensor(s).prod())].reshape(s)
         for k, s in o.shapes().items()}
    ids = torch.as_tensor(ids, dtype=torch.long) - 1
    N = ids.shape[0]
    H = c.H
    h = torch.zeros(N, H, dtype=torch.float64)
    cc = torch.zeros(N, H, dtype=torch.float64)
    for t in range(c.T):
        x = torch.cat([P["E_t"][ids[:, t, 0]], P["E_e"][ids[:, t, 1]], P["E_r"][ids[:, t, 2]]], dim=1)
        if c.cell == 0:
            G = x @ P["Wx"].T + P["bx"] + h @ P["Wh"].T
            i, g, f, og = torch.sigmoid(G[:, :H]), torch.tanh(G[:, H:2 * H]), torch.sigmoid(G[:, 2 * H:3 * H]), torch.sigmoid(G[:, 3 * H:])
            cc = i * g + f * cc
            h = og * torch.tanh(cc)
        else:
            v = torch.relu(x @ P["Wx"].T + P["bx"] + h @ P["Wh"].T + P["bh"])
            mask = (x.abs().sum(dim=1, keepdim=True) != 0).to(v.dtype)
            h = v * mask
    z = h
    if c.head_layers == 2:
        z = torch.relu(h @ P["W1"].T + P["b1"])
    s = z @ P["Wo"].T + P["bo"]
    outs = []
    ig = c.inv_gamma
    for k in range(len(seg) - 1):
        sk = s[int(seg[k]):int(seg[k + 1])]
        if c.pool == 0:
            outs.append(torch.logsumexp(sk * ig, dim=0) / ig)
        elif c.pool == 1:
            outs.append(sk.max(dim=0).values)
        else:
            kk = min(c.topk, sk.shape[0])
            outs.append(sk.topk(kk, dim=0).values.mean(dim=0))
    lse = torch.stack(outs)
    return s, lse, torch.sigmoid(lse[:, class_id - 1])


def bce(p, y):
    eps = 1e-12
    return -(y * torch.log(p + eps) + (1 - y) * torch.log(1 - p + eps)).mean()
