"""Shared helpers for the parity tests."""
import torch


def rel(a, b):
    a, b = torch.as_tensor(a).detach().cpu().double(), torch.as_tensor(b).detach().cpu().double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def align_selection(got_idx, want_idx, want_sdf, tie_tol=2e-7):
    """Selections are lists of lattice indices ordered by ascending |sdf|.  Two fp32 evaluations of the same
    network (even upstream vs upstream on another BLAS) differ by ~1e-8, which can swap neighbours whose |sdf|
    are closer than that, so equality is required as SETS, and the ORDER may differ only between entries whose
    reference |sdf| differ by less than `tie_tol`.  Returns, per sample, the permutation `perm` such that
    got_idx[b][perm[b]] == want_idx[b] (use it to align per-point outputs before comparing them)."""
    got_idx, want_idx = torch.as_tensor(got_idx).cpu().long(), torch.as_tensor(want_idx).cpu().long()
    want_abs = torch.as_tensor(want_sdf).cpu().reshape(want_idx.shape).abs()
    perms = []
    for b in range(want_idx.shape[0]):
        g, w = got_idx[b].tolist(), want_idx[b].tolist()
        assert sorted(g) == sorted(w), "selected point SETS differ for sample %d" % b
        pos = {v: j for j, v in enumerate(g)}
        perm = torch.tensor([pos[v] for v in w])
        moved = (perm != torch.arange(len(w))).nonzero().flatten()
        for j in moved.tolist():
            # the entry that sits at rank j on our side must be a near-tie of the reference's rank-j entry
            other = w.index(g[j])
            assert abs(float(want_abs[b, j]) - float(want_abs[b, other])) < tie_tol, \
                (b, j, other, float(want_abs[b, j]), float(want_abs[b, other]))
        perms.append(perm)
    return perms


def aligned(t, perms, dim=1):
    """Reorder the per-point dimension of `t` (B, P, ...) with the permutations from align_selection."""
    t = torch.as_tensor(t).detach().cpu()
    return torch.stack([t[b].index_select(dim - 1, perms[b]) for b in range(t.shape[0])])
