"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def sample_idx(numel, count=24):
    return (np.arange(count, dtype=np.int64) * 2654435761 % max(numel, 1)).astype(np.int64)


def summarize(t, count=24):
    f = t.detach().reshape(-1).double().cpu()
    idx = sample_idx(f.numel(), count)
    return np.concatenate([[f.norm().item(), f.sum().item()], f[torch.from_numpy(idx)].numpy()])


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def max_abs(a, b):
    return float((torch.as_tensor(a).double().cpu() - torch.as_tensor(b).double().cpu()).abs().max())


def assert_summary_close(actual, desired, tol_norm=2e-3, tol_samp=1e-2, tol_scalar=6e-2, what="", abs_slack=0.0, scalar_abs=0.0):
    """Compare stacks of summarize() rows.  The raw `sum` column is ignored (cancellation noise); the
    L2 norm must agree to tol_norm, the strided samples to tol_samp x max|sample|.  Rows whose samples
    are all identical are 1-element tensors (the scalar norm gains/biases): their gradient is one global,
    cancellation-heavy sum, so they get tol_scalar.  abs_slack: absolute allowance on norm and samples (Adam moves an
    element whose gradient is ~0 by +-lr depending on rounding noise: 2*lr per flipped element).  scalar_abs: absolute
    allowance for the scalar rows only (a scalar gradient that cancels to ~1e-3 of its siblings' magnitude carries the
    rounding noise of the whole sum)."""
    actual = np.asarray(actual)
    desired = np.asarray(desired)
    assert actual.shape == desired.shape
    for i in range(desired.shape[0]):
        a, e = actual[i], desired[i]
        scalar = np.all(e[2:] == e[2])
        tn = tol_scalar if scalar else tol_norm
        ts = tol_scalar if scalar else tol_samp
        if scalar:
            abs_slack_row = abs_slack + scalar_abs
        else:
            abs_slack_row = abs_slack
        assert abs(a[0] - e[0]) <= tn * abs(e[0]) + abs_slack_row + 1e-12, "%s row %d norm %g vs %g" % (what, i, a[0], e[0])
        scale = np.abs(e[2:]).max() + 1e-30
        err = np.abs(a[2:] - e[2:]).max()
        assert err <= ts * scale + abs_slack_row, "%s row %d sample err %g scale %g" % (what, i, err, scale)
