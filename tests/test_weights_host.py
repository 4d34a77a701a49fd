"""Class-balanced weights (host half): cvcs_b200.loss.class_weights_from_counts and the oracle's
restatement against dataset.Loader.get_class_weights run unmodified on seeded random class counts
(tests/golden/weight_cases.npz, dataset.py:360-384) and on the tiny on-disk dataset of dataset_cases.npz."""
import numpy as np
import pytest
import torch

from oracle import torch_path


@pytest.mark.parametrize("i", range(16))
@pytest.mark.parametrize("ib", [False, True])
def test_class_weights_are_the_references(golden, i, ib):
    from cvcs_b200.loss import class_weights_from_counts
    g = golden("weight_cases")
    counts = torch.from_numpy(g[f"w{i}.counts"])
    ref = g[f"w{i}.ib{int(ib)}"]
    w = class_weights_from_counts(counts, ib)
    assert w.numpy().dtype == ref.dtype and np.array_equal(w.numpy(), ref)      # all-empty counts give int64 zeros, as there
    # K4 hands the counts over as int64: same weights
    assert np.array_equal(class_weights_from_counts(counts.to(torch.int64), ib).numpy(), ref)
    assert np.array_equal(torch_path.class_weights(counts, ib).numpy(), ref)


def test_class_weights_of_the_golden_dataset(golden):
    from cvcs_b200.loss import class_weights_from_counts
    g = golden("dataset_cases")
    counts = torch.from_numpy(g["counts"])
    assert np.array_equal(class_weights_from_counts(counts, False).numpy(), g["weights_ib0"])
    assert np.array_equal(class_weights_from_counts(counts, True).numpy(), g["weights_ib1"])
