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


def test_group_schedule_covers_every_tile_once_and_keeps_the_sharers_together():
    """Group mode of tile_at (csrc/gemm_tcgen05.cu, sched_shift == -2; A/B knob tgb_gemm_set_schedule(3)): the grid is
    C = (clusters // NB) * NB clusters, round r of cluster c is the tile r*C + NB*(c // NB) + (c % NB + r) % NB.  Every tile
    is computed exactly once, the NB clusters of a group always share one A row block, and every cluster meets the cheap
    last N tile once every NB rounds."""
    for clusters, nb, mb in [(74, 6, 514), (74, 17, 514), (74, 11, 514), (9, 3, 10)]:
        used = clusters // nb * nb
        tiles = nb * mb
        seen = [0] * tiles
        last = [0] * used
        rounds = [0] * used
        for c in range(used):
            r = 0
            while True:
                t = r * used + nb * (c // nb) + (c % nb + r) % nb
                if t >= tiles:
                    break
                seen[t] += 1
                assert t // nb == r * (used // nb) + c // nb           # the group's row block of this round
                last[c] += t % nb == nb - 1
                rounds[c] += 1
                r += 1
        assert all(v == 1 for v in seen), (clusters, nb, mb)
        assert all(abs(l - r / nb) <= 1 for l, r in zip(last, rounds))


def _slot_columns(n, bn, pw):
    """Columns summed into every statistics slot by the fp32-residual epilogue with tile width `bn` and piece width `pw`
    (csrc/gemm_tcgen05.cu): warp `sub` of a lane quarter walks the pieces sub, sub + 2, ... of a tile, a slot is
    (128-column block of N, sub), flushed when the walk leaves the block.  Returns {slot: [columns in summation order]}."""
    slots = {}
    for col0 in range(0, n, bn):
        for sub in range(2):
            j = sub
            while j < bn // pw and col0 + j * pw < n:
                col = col0 + j * pw
                slots.setdefault(((col >> 7), sub), []).extend(c for c in range(col, col + pw) if c < n)
                j += 2
    return slots


def test_statistics_slots_do_not_depend_on_the_tile_width():
    """The LayerNorm-fold statistics are batch-invariant because the columns a slot sums, and their order, depend on N and
    the piece width only -- not on whether a pass took 256- or 128-wide tiles."""
    for n in (1408, 768, 1024, 136, 272):
        for pw in (16, 32):
            wide, narrow = _slot_columns(n, 256, pw), _slot_columns(n, 128, pw)
            assert wide == narrow, (n, pw)
            assert sorted(c for cols in wide.values() for c in cols) == list(range(n))      # every column exactly once
            assert len(wide) <= ((n + 127) // 128) * 2                                      # tgb_gemm_ln_stats_slots
