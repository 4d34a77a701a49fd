"""Loader / IterableChunk drop-in (cvcs_b200/dataset.py) against fixtures produced by the UNMODIFIED
reference ``dataset.Loader`` (tests/golden/make_golden.py::dataset_cases) and against the oracle."""
import os
import random

import numpy as np
import pytest
import torch

from oracle import c_oracle

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _scenes(g):
    imgs = [torch.from_numpy(g[f"scene{i}.image"]) for i in (0, 1)]
    labs = [torch.from_numpy(g[f"scene{i}.label"]) for i in (0, 1)]
    cols = [torch.from_numpy(g[f"scene{i}.color"]) for i in (0, 1)]
    return imgs, labs, cols


@pytest.mark.parametrize("shift", [False, True])
def test_loader_matches_reference_chunk(golden, shift):
    """Same seed -> same shuffled crop order, same random shifts, same tiles, bit for bit."""
    from cvcs_b200.dataset import ArrayScenes, Loader
    g = golden("dataset_cases")
    imgs, labs, cols = _scenes(g)
    tag = f"shift{int(shift)}"
    random.seed(1234)
    L = Loader(ArrayScenes(imgs, labs, cols), chunk_size=2, random_shift=shift, patch_size=224, load_context=False,
               load_color_mask=True, device=DEV)
    chunk = L.get_iterable_chunk(0)
    assert len(L) == int(g[f"{tag}.len"]) and chunk.tpi == int(g[f"{tag}.tpi"])
    assert chunk.chunk_crops == g[f"{tag}.chunk_crops"].tolist()
    assert len(chunk.patches) == len(chunk.chunk_crops)
    for k, (patch, index_mask, color_mask, context) in enumerate(chunk):
        assert patch.dtype == torch.uint8 and index_mask.shape == (224, 224)
        assert np.array_equal(patch.cpu().numpy(), g[f"{tag}.patches"][k])
        assert np.array_equal(index_mask.cpu().numpy(), g[f"{tag}.index_masks"][k])
        assert np.array_equal(color_mask.cpu().numpy(), g[f"{tag}.color_masks"][k])
        assert context.tolist() == [0]
    bi, bl, bc = chunk.batch
    assert np.array_equal(bi.cpu().numpy(), g[f"{tag}.patches"])
    # the training loop's DataLoader over the chunk (train.py:113) collates GPU tensors
    dl = torch.utils.data.DataLoader(chunk, batch_size=2)
    image, index_mask, color_mask, context = next(iter(dl))
    assert image.is_cuda and image.shape == (2, 4, 224, 224) and index_mask.shape == (2, 224, 224)
    assert np.array_equal(image.cpu().numpy(), g[f"{tag}.patches"][:2])


def test_loader_directory_tree_and_class_weights(golden, tmp_path):
    """The GID-15 directory layout (decoded with PIL), and get_class_weights / counts."""
    from PIL import Image
    from cvcs_b200.dataset import Loader
    g = golden("dataset_cases")
    root = str(tmp_path)
    for sub in ("Image__8bit_NirRGB", "Annotation__index", "Annotation__color"):
        os.makedirs(os.path.join(root, sub))
    for si, stem in enumerate(["scene_a", "scene_b"]):
        Image.fromarray(g[f"scene{si}.image"].transpose(1, 2, 0).copy(), "RGBA").save(
            os.path.join(root, "Image__8bit_NirRGB", stem + ".png"))
        Image.fromarray(g[f"scene{si}.label"], "L").save(os.path.join(root, "Annotation__index", stem + "_15label.png"))
        Image.fromarray(g[f"scene{si}.color"].transpose(1, 2, 0).copy(), "RGB").save(
            os.path.join(root, "Annotation__color", stem + "_15label.tif"))
    random.seed(1234)
    L = Loader(root, chunk_size=2, random_shift=True, patch_size=224, load_context=False, load_color_mask=True, device=DEV)
    assert L.image_shape == [230, 460] and L.tpi == 2
    chunk = L.get_iterable_chunk(0)
    assert np.array_equal(torch.stack([t[0] for t in chunk.patches]).cpu().numpy(), g["shift1.patches"])
    assert np.array_equal(torch.stack([t[2] for t in chunk.patches]).cpu().numpy(), g["shift1.color_masks"])
    L1 = Loader(root, chunk_size=1, patch_size=224, load_context=False, load_color_mask=False, device=DEV)
    w0 = L1.get_class_weights(16, False)
    assert np.array_equal(L1.count.numpy(), g["counts"]) and np.array_equal(w0.numpy(), g["weights_ib0"])
    L2 = Loader(root, chunk_size=1, patch_size=224, load_context=False, load_color_mask=False, device=DEV)
    assert np.array_equal(L2.get_class_weights(16, True).numpy(), g["weights_ib1"])
    with pytest.raises(AssertionError, match="Patch size"):
        Loader(root, patch_size=1024, device=DEV)
    with pytest.raises(AssertionError, match="not divisible"):
        Loader(root, chunk_size=3, patch_size=224, device=DEV)


def test_float_tiles_fused_cast_normalize_and_hist(golden):
    from cvcs_b200.dataset import ArrayScenes, Loader
    g = golden("dataset_cases")
    imgs, labs, cols = _scenes(g)
    random.seed(7)
    L = Loader(ArrayScenes(imgs, labs), chunk_size=2, random_shift=True, patch_size=224, load_context=False,
               load_color_mask=False, device=DEV)
    chunk = L.get_iterable_chunk(0)
    bi, bl, bc = chunk.batch
    assert bc is None
    x, y = chunk.float_tiles()
    assert torch.equal(x, bi.type(torch.float32)) and torch.equal(y, bl)          # train.py:121
    mean = torch.tensor([10.0, 20.0, 30.0, 40.0], device=DEV)
    std = torch.tensor([3.0, 5.0, 7.0, 11.0], device=DEV)
    hist = torch.zeros(18, dtype=torch.int64, device=DEV)
    xn, _ = chunk.float_tiles(mean, std, hist=hist, hist_classes=16)
    for k in range(len(chunk.chunk_crops)):
        s = chunk.tile_scene[k]
        yx = np.array([chunk.tile_yx[k]], dtype=np.int32)
        ref, ref_lab = c_oracle.tile(imgs[s].numpy(), yx, 224, 224, mean.cpu().numpy(), std.cpu().numpy(),
                                     labels=labs[s].numpy())
        assert np.array_equal(xn[k].cpu().numpy(), ref[0]) and np.array_equal(y[k].cpu().numpy(), ref_lab[0])
    assert np.array_equal(hist.cpu().numpy(), c_oracle.label_hist(y.cpu().numpy(), 16))
    xb, _ = chunk.float_tiles(mean, std, dtype=torch.bfloat16)
    assert torch.equal(xb, xn.to(torch.bfloat16))


def test_loader_shuffle_specify_and_context():
    from cvcs_b200.dataset import ArrayScenes, Loader
    gen = torch.Generator().manual_seed(0)
    imgs = [torch.randint(0, 256, (3, 96, 128), generator=gen, dtype=torch.uint8) for _ in range(4)]
    labs = [torch.randint(0, 16, (96, 128), generator=gen, dtype=torch.uint8) for _ in range(4)]
    random.seed(3)
    L = Loader(ArrayScenes(imgs, labs), chunk_size=2, patch_size=32, load_context=True, load_color_mask=False,
               device=DEV, strict_patch_size=False)
    assert len(L) == 2 and L.tpi == 12
    ref_idxs = list(range(4))
    random.seed(3)
    random.shuffle(ref_idxs)
    random.seed(3)
    L.shuffle()
    assert L.idxs == ref_idxs and L.chunks == [ref_idxs[:2], ref_idxs[2:]]
    assert L.get_chunk(1) == [L.images[i] for i in ref_idxs[2:]]
    L.specify([0, 1])
    assert len(L) == 1
    chunk = L.get_iterable_chunk(0, random_tps=[(48, 0.25)])
    assert len(chunk.patches) == 24 + 6
    for patch, index_mask, color_mask, context in chunk:
        assert patch.shape == (3, 32, 32) and index_mask.shape == (32, 32) and context.shape == (3, 32, 32)
        assert color_mask.tolist() == [0]
    # the context of every regular tile = _get_context of its scene at its origin (dataset.py:154), byte for byte
    L.specify([0, 1])
    chunk = L.get_iterable_chunk(0)
    scenes = [imgs[i].numpy() for i in L.chunks[0]]
    for k, (patch, _, _, context) in enumerate(chunk):
        s, (tly, tlx) = chunk.tile_scene[k], chunk.tile_yx[k]
        want = c_oracle.context(scenes[s], np.array([[tly, tlx]], dtype=np.int32), 32)[0]
        assert np.array_equal(context.cpu().numpy(), want)
        assert np.array_equal(patch.cpu().numpy(), scenes[s][:, tly:tly + 32, tlx:tlx + 32])
