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


# ----------------------------------------------------------------------------------------------------
# training-step parity helpers (fixture tests/golden/train_*.npz, oracle.model_train, hoisdf_b200.train)
# ----------------------------------------------------------------------------------------------------
GRAD_PROBES = 24


def grad_summary(g):
    """abs-max, L2 norm and GRAD_PROBES evenly strided entries of a gradient (the layout oracle/make_golden.py stores)."""
    f = torch.as_tensor(g).detach().cpu().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(GRAD_PROBES, f.numel())).long()
    return torch.cat([torch.tensor([float(f.abs().max()), float(f.norm())], dtype=torch.float64), f[idx]])


def param_group(name):
    """Parameter group of a state-dict key: the sub-network a gradient tolerance is taken relative to (a parameter whose
    true gradient is ~0 -- a convolution bias in front of a BatchNorm, the first decoder layer's q/k projection of an
    all-zero target -- cannot be compared relative to itself)."""
    parts = name.split(".")
    if parts[0] in ("hand_transformer", "obj_transformer"):
        return ".".join(parts[:2])
    return parts[0]


def group_scales(ref):
    """{group: max |gradient| over the group's tensors} from {name: gradient summary or gradient tensor}."""
    scales = {}
    for n, g in ref.items():
        m = float(torch.as_tensor(g).detach().abs().max()) if torch.as_tensor(g).dim() != 1 or len(g) != GRAD_PROBES + 2 \
            else float(g[0])
        scales[param_group(n)] = max(scales.get(param_group(n), 0.0), m)
    return scales


def oracle_train_step(seed, arch, batch, ph, po, device="cpu", dtype=torch.float32):
    """Run oracle.model_train + backward on the seeded synthetic training batch.  Returns (model_out, weighted parts,
    total, {name: grad}) -- the reference gradients of main/train.py:111-131 (dropout 0, zero jitter).  dtype float64
    evaluates the same (upstream-pinned) formulas in double precision: the yardstick two fp32 implementations are
    measured against (smooth-L1 on millimetre residuals and ReLU / threshold decisions make single fp32 gradients jump)."""
    from hoisdf_b200 import synthetic as syn
    from oracle import hoisdf_oracle as O
    cv = lambda v: (v.clone().to(dtype) if v.is_floating_point() else v.clone()).to(device)  # noqa: E731
    sd = syn.full_state_dict(seed, arch)
    p = {k: cv(v) for k, v in sd.items()}
    names = [k for k, v in p.items() if v.is_floating_point() and "running_" not in k and "th_" not in k
             and "num_batches" not in k and "coord_change" not in k]
    for n in names:
        p[n].requires_grad_(True)
    mv = lambda d: {k: cv(v) for k, v in d.items()}  # noqa: E731
    inputs, targets = syn.train_extras(seed, batch, ph, po)
    out = O.model_train(p, cv(syn.image_batch(seed, batch)), mv(inputs), mv(targets),
                        mv(syn.camera_meta(seed, batch)), O.default_cfg(num_samp_hand=ph, num_samp_obj=po), arch)
    total, parts = O.train_total_loss(out)
    total.backward()
    grads = {n: p[n].grad for n in names if p[n].grad is not None}
    return out, parts, total, grads
