"""Host-side logic that needs no GPU: the LayerNorm-fold weight algebra and the persistent tile schedule."""
import torch

from videotgb_b200.modeling import Blip2VisionModel, _LayerNorm, _Linear


def test_layernorm_fold_algebra():
    """`_fold_ln` must satisfy LN(x) W^T + b == rstd * (x W'^T - mean * colsum(W')) + bias' (up to the bf16 rounding of W',
    which colsum is taken over so that the mean term cancels what the tensor cores add)."""
    g = torch.Generator().manual_seed(2)
    k, n, m = 96, 40, 17
    ln, lin = _LayerNorm(k, 1e-6), _Linear(k, n)
    with torch.no_grad():
        ln.weight.copy_(1 + 0.3 * torch.randn(k, generator=g))
        ln.bias.copy_(0.2 * torch.randn(k, generator=g))
        lin.weight.copy_(torch.randn(n, k, generator=g) * 0.1)
        lin.bias.copy_(torch.randn(n, generator=g))
    wf, colsum, bias = Blip2VisionModel._fold_ln(ln, lin)
    assert wf.dtype == torch.bfloat16 and wf.shape == (n, k) and colsum.shape == (n,) and bias.shape == (n,)
    assert torch.equal(colsum, wf.float().sum(dim=1))
    x = (torch.randn(m, k, generator=g) * 2 + 0.7).double()
    mean = x.mean(dim=1, keepdim=True)
    rstd = 1.0 / torch.sqrt(x.var(dim=1, unbiased=False, keepdim=True) + 1e-6)
    folded = rstd * (x @ wf.double().t() - mean * colsum.double()) + bias.double()
    # the same computation with the un-rounded folded weight is exactly LayerNorm followed by the linear layer
    w_exact = (lin.weight.detach() * ln.weight.detach()[None, :]).double()
    exact = rstd * (x @ w_exact.t() - mean * w_exact.sum(dim=1)) + bias.double()
    normed = torch.nn.functional.layer_norm(x, (k,), ln.weight.detach().double(), ln.bias.detach().double(), 1e-6)
    want = normed @ lin.weight.detach().double().t() + lin.bias.detach().double()
    assert (exact - want).abs().max().item() < 1e-5
    assert (folded - want).abs().max().item() < 2 ** -7 * want.abs().max().item()      # only W' -> bf16 separates them


def _walk(clusters, nb, mb, shift):
    tiles = nb * mb
    seen = [0] * tiles
    cost = [0.0] * clusters
    for c in range(clusters):
        rnd = 0
        while True:
            t = rnd * clusters + (c + rnd * shift) % clusters
            if t >= tiles:
                break
            seen[t] += 1
            cost[c] += 0.65 if t % nb == nb - 1 else 1.0
            rnd += 1
    return seen, cost


def test_persistent_schedule_covers_every_tile_once_and_balances():
    """tile_at(round) = round*C + (c + round*shift) mod C (csrc/gemm_tcgen05.cu): whatever the shift, every tile is computed
    exactly once; the searched shift (balanced_shift, restated here) brings the busiest cluster close to the mean where the
    plain round-robin leaves it 7 % above (N = 1408: six N blocks, the last one half wide)."""
    for clusters, nb, mb in [(74, 6, 514), (74, 17, 514), (37, 6, 40), (5, 3, 4)]:
        for shift in (0, 1, 5, clusters - 1):
            seen, _ = _walk(clusters, nb, mb, shift)
            assert all(v == 1 for v in seen), (clusters, nb, mb, shift)
    clusters, nb, mb = 74, 6, 514
    _, rr = _walk(clusters, nb, mb, 0)
    best = min(range(clusters), key=lambda sh: max(_walk(clusters, nb, mb, sh)[1]))
    _, bal = _walk(clusters, nb, mb, best)
    mean = sum(rr) / clusters
    assert max(rr) > 1.06 * mean and max(bal) < 1.01 * mean


def test_row_block_schedule_covers_every_tile_once_and_keeps_row_blocks_private():
    """Row-block mode of tile_at (csrc/gemm_tcgen05.cu, sched_shift < 0): round r of cluster c is the tile
    (c + (r // NB) * C) * NB + (r % NB + c % NB) % NB.  Every tile is computed exactly once and every A row block by
    exactly one cluster; for config 2 (514 row blocks, 74 clusters) the busiest cluster does the same work as under the
    searched tile-granular schedule."""
    for clusters, nb, mb in [(74, 6, 514), (74, 17, 514), (74, 24, 514), (74, 6, 74), (3, 2, 7)]:
        tiles = nb * mb
        seen = [0] * tiles
        owner = {}
        cost = [0.0] * clusters
        for c in range(clusters):
            r = 0
            while True:
                t = (c + (r // nb) * clusters) * nb + (r % nb + c % nb) % nb
                if t >= tiles:
                    break
                seen[t] += 1
                owner.setdefault(t // nb, set()).add(c)
                cost[c] += 0.65 if t % nb == nb - 1 else 1.0
                r += 1
        assert all(v == 1 for v in seen), (clusters, nb, mb)
        assert all(len(v) == 1 for v in owner.values())
    clusters, nb, mb = 74, 6, 514
    best = min(range(clusters), key=lambda sh: max(_walk(clusters, nb, mb, sh)[1]))
    _, bal = _walk(clusters, nb, mb, best)
    rows = [(7 if c < mb - 6 * clusters else 6) * (nb - 1 + 0.65) for c in range(clusters)]
    assert abs(max(rows) - max(bal)) < 1e-6 * max(bal) + 0.7      # within one half tile of the searched schedule
