"""Index builder for benchmarks and large tests (SETUP tooling, never on the timed search path).

Produces what GammaIVFPQIndex::Indexing + Add produce on the CPU (gamma_index_ivfpq.cc:272-512):
coarse centroids (k-means, niter=10 as `cp.niter = 10`, :175), PQ codebooks trained on coarse
residuals (by_residual = true, :179), and for every vector its list number and M-byte code.
It runs on the GPU with plain torch ops because training a 16k-centroid quantiser on the host
takes hours; the hot path itself never imports this module.  The same trained state and codes
are handed to BOTH engines (device mirror via the C-ABI, reference CPU engine via
oracle.ref.RefIndex.set_trained / inject_postings), so they search an identical index.
"""
import numpy as np
import torch


def _assign(x, c, c_norm, chunk=65536):
    """argmin_j |x_i - c_j|^2 for rows of x (n,d) against c (k,d)."""
    out = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
    chunk = max(1024, min(chunk, (1 << 29) // max(1, c.shape[0])))  # keep the distance block around 2 GB
    for s in range(0, x.shape[0], chunk):
        xs = x[s:s + chunk]
        d = c_norm[None, :] - 2.0 * (xs @ c.t())
        out[s:s + chunk] = d.argmin(dim=1)
    return out


def kmeans(x, k, niter, seed):
    """Lloyd k-means, faiss-style: random-sample init, empty clusters re-seeded by splitting big ones."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    n, d = x.shape
    perm = torch.randperm(n, generator=g)[:k].to(x.device)
    c = x[perm].clone()
    for _ in range(niter):
        a = _assign(x, c, (c * c).sum(1))
        sums = torch.zeros_like(c).index_add_(0, a, x)
        cnt = torch.bincount(a, minlength=k).to(x.dtype)
        nz = cnt > 0
        c[nz] = sums[nz] / cnt[nz, None]
        empty = (~nz).nonzero().flatten()
        if empty.numel():
            big = torch.argsort(cnt, descending=True)[:empty.numel()]
            c[empty] = c[big] * (1 + 1e-4)
            c[big] = c[big] * (1 - 1e-4)
    return c


def train_pq(res, M, niter, seed):
    """res (n,d) residuals -> codebooks (M,256,dsub), all sub-quantisers batched."""
    n, d = res.shape
    dsub = d // M
    x = res.view(n, M, dsub).permute(1, 0, 2).contiguous()  # (M,n,dsub)
    g = torch.Generator(device="cpu").manual_seed(seed)
    perm = torch.randperm(n, generator=g)[:256].to(res.device)
    c = x[:, perm, :].clone()  # (M,256,dsub)
    for _ in range(niter):
        dist = (c * c).sum(2)[:, None, :] - 2.0 * torch.bmm(x, c.transpose(1, 2))  # (M,n,256)
        a = dist.argmin(dim=2)  # (M,n)
        onehot_cnt = torch.zeros(M, 256, device=res.device, dtype=res.dtype)
        onehot_cnt.scatter_add_(1, a, torch.ones_like(a, dtype=res.dtype))
        sums = torch.zeros_like(c)
        sums.scatter_add_(1, a[:, :, None].expand(-1, -1, dsub), x)
        nz = onehot_cnt > 0
        c = torch.where(nz[:, :, None], sums / onehot_cnt.clamp(min=1)[:, :, None], c)
    return c


def encode(x, coarse, pq, chunk=1 << 18):
    """list number + PQ code of every row of x (by_residual encoding)."""
    n, d = x.shape
    M, _, dsub = pq.shape
    list_no = torch.empty(n, dtype=torch.int32, device=x.device)
    codes = torch.empty(n, M, dtype=torch.uint8, device=x.device)
    cn = (coarse * coarse).sum(1)
    pn = (pq * pq).sum(2)  # (M,256)
    for s in range(0, n, chunk):
        xs = x[s:s + chunk]
        a = _assign(xs, coarse, cn, chunk=32768)
        r = (xs - coarse[a]).view(-1, M, dsub).permute(1, 0, 2)  # (M,b,dsub)
        dist = pn[:, None, :] - 2.0 * torch.bmm(r, pq.transpose(1, 2))
        codes[s:s + chunk] = dist.argmin(dim=2).t().to(torch.uint8)
        list_no[s:s + chunk] = a.to(torch.int32)
    return list_no, codes


def build_ivfpq(xb, nlist, M, device="cuda", train_n=None, seed=1234, coarse_iters=10, pq_iters=25,
                upload_chunk=1 << 20):
    """xb: (N,d) float32 numpy.  Returns numpy (coarse[nlist,d], pq[M,256,dsub], list_no[N] i32, codes[N,M] u8).

    Training set = the first train_n vectors, train_n clamped to [39*nlist, 256*nlist] like the
    reference (gamma_index_ivfpq.cc:280-296); the lower bound keeps setup short."""
    N, d = xb.shape
    if train_n is None:
        train_n = 39 * nlist
    train_n = int(min(N, max(39 * nlist, min(train_n, 256 * nlist))))
    dev = torch.device(device)
    xt = torch.from_numpy(xb[:train_n]).to(dev)
    coarse = kmeans(xt, nlist, coarse_iters, seed)
    a = _assign(xt, coarse, (coarse * coarse).sum(1))
    res = xt - coarse[a]
    pq = train_pq(res, M, pq_iters, seed + 1)
    del xt, res
    list_no = np.empty(N, np.int32)
    codes = np.empty((N, M), np.uint8)
    for s in range(0, N, upload_chunk):
        xs = torch.from_numpy(xb[s:s + upload_chunk]).to(dev)
        ln, cd = encode(xs, coarse, pq)
        list_no[s:s + upload_chunk] = ln.cpu().numpy()
        codes[s:s + upload_chunk] = cd.cpu().numpy()
    return coarse.cpu().numpy(), pq.cpu().numpy(), list_no, codes
