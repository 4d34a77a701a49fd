"""Border-corrected inference drop-in (cvcs_b200/inference.py) against the reference's own semantics
re-enacted with torch / torchvision on the host (dataset.py:18-23,70-96; utils.py:145-171; inference.py:40-57)."""
import numpy as np
import pytest
import torch
import torchvision.transforms as T
import torchvision.transforms.v2 as v2

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


class StubNet(torch.nn.Module):
    """A deterministic 'segmenter': 3x3 box filter of the bands followed by a fixed 1x1 projection, so that
    border pixels really depend on the padding the patch was cut with."""
    requires_context = False
    returns_logits = True

    def __init__(self, cb, classes):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        self.register_buffer("proj", torch.randn(classes, cb, generator=g))

    def forward(self, x, context=None):
        x = torch.nn.functional.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=True)
        return torch.einsum("kc,bchw->bkhw", self.proj.to(x.dtype), x)


@pytest.mark.parametrize("p,bc", [(32, None), (32, 40), (32, 35), (28, 39)])
def test_inference_scene_matches_reference_semantics(p, bc):
    from cvcs_b200.inference import inference_scene
    g = torch.Generator().manual_seed(p + (bc or 0))
    Cb, H, W, C = 4, 100, 140, 6
    scene = torch.randint(0, 256, (Cb, H, W), generator=g, dtype=torch.uint8)
    net = StubNet(Cb, C).double()         # fp64 on both sides: the comparison is about geometry, not rounding
    rows, cols = H // p, W // p
    # --- the reference's procedure on the host, tile by tile
    ref = torch.zeros((rows * p, cols * p), dtype=torch.uint8)
    crop = T.CenterCrop(p)
    for idx in range(rows * cols):
        tly, tlx = (idx // cols) * p, (idx % cols) * p
        if bc:
            margin = bc - p
            patch = v2.functional.crop(scene, tly - margin, tlx - margin, bc, bc)     # _get_padded_patch
            output = crop(net(patch.unsqueeze(0).double()))
        else:
            output = net(v2.functional.crop(scene, tly, tlx, p, p).unsqueeze(0).double())
        pred = torch.argmax(output.squeeze().permute(1, 2, 0), dim=2)               # utils.py:158
        ref[tly:tly + p, tlx:tlx + p] = pred.to(torch.uint8)                        # inference.py:40-57 re-assembly

    class Net64(StubNet):
        def forward(self, x, context=None):
            return super().forward(x.double()).float()
    net_gpu = Net64(Cb, C).double().to(DEV)
    out = inference_scene(net_gpu, scene.to(DEV), p, bc, batch_size=5)
    assert out.shape == ref.shape
    assert np.array_equal(out.cpu().numpy(), ref.numpy())
    # `range` option: only some tiles
    out2 = inference_scene(net_gpu, scene.to(DEV), p, bc, batch_size=4, indexes=[1, 2])
    assert np.array_equal(out2[:p, p:3 * p].cpu().numpy(), ref[:p, p:3 * p].numpy()) and int(out2[p:].sum()) == 0


def test_gid15_item_layout_and_padded_patch(golden):
    from cvcs_b200.dataset import ArrayScenes
    from cvcs_b200.inference import GID15
    g = golden("dataset_cases")
    imgs = [torch.from_numpy(g[f"scene{i}.image"]) for i in (0, 1)]
    labs = [torch.from_numpy(g[f"scene{i}.label"]) for i in (0, 1)]
    ds = GID15(ArrayScenes(imgs, labs), (224, 224), border_correction=256, device=DEV)
    assert len(ds) == 2 * 2 and ds.tiles_in_img_shape == (1, 2)
    tif, mask, context, padded = ds[3]                      # scene 1, tile 1 -> (0, 224)
    assert np.array_equal(tif.cpu().numpy(), imgs[1][:, :224, 224:448].numpy())
    assert np.array_equal(mask.cpu().numpy()[0], labs[1][:224, 224:448].numpy())
    ref_pad = v2.functional.crop(imgs[1], 0 - 32, 224 - 32, 256, 256)               # dataset.py:18-23
    assert np.array_equal(padded.cpu().numpy(), ref_pad.numpy())
    assert context.shape == (4, 224, 224)
    with pytest.raises(TypeError):
        GID15(ArrayScenes(imgs, labs), (224, 224), random_shift=True, device=DEV)[0]
