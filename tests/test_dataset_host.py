"""Host-side tile index arithmetic of the Loader drop-in (no GPU): against the reference-generated
fixtures and the oracle's restatement of dataset.py:125,136-140."""
import numpy as np

from oracle import torch_path


def test_tile_origin_matches_oracle_and_reference_order(golden):
    from cvcs_b200 import dataset
    g = golden("dataset_cases")
    assert dataset.tiles_in_image([230, 460], 224) == (1, 2)
    assert dataset.tiles_per_image([230, 460], 224) == int(g["shift0.tpi"]) == 2
    assert dataset.tiles_per_image([6800, 7200], 224) == 960                    # utils.py:75
    assert dataset.tiles_in_image([10000, 10000], 1024) == (9, 9)               # cfg4: 81 tiles / scene
    for H, W, p in [(6800, 7200, 224), (10000, 10000, 1024), (230, 460, 224), (100, 37, 8)]:
        rows, cols = dataset.tiles_in_image([H, W], p)
        assert (rows, cols) == tuple(torch_path.tiles_in_image(H, W, p))
        tpi = rows * cols
        for x in list(range(min(3 * tpi, 200))) + [5 * tpi - 1]:
            assert dataset.tile_origin(x, tpi, cols, p) == tuple(torch_path.tile_top_left(x, tpi, cols, p))
    # every tile lies inside the scene; the remainder strip is never touched
    rows, cols = dataset.tiles_in_image([10000, 10000], 1024)
    ys = {dataset.tile_origin(x, rows * cols, cols, 1024)[1] for x in range(rows * cols)}
    assert max(ys) + 1024 <= 10000 and len(ys) == rows
