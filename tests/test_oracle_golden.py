"""CPU: the oracle (oracle/adabins_oracle.py) replays the golden vectors produced by the REFERENCE modules
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU parity tests then compare against it."""
import numpy as np
import pytest
import torch

from oracle import adabins_oracle as oracle
from mde_biological_vision_systems_b200 import synthetic

from helpers import INST_MODES, SEM_MODES, digest, inst_labels, load_table, make_model, sem_labels

SEM_TABLE = {
    "glove": "ade20k_150_classes_glove_840b_300d_embeddings.npy",
    "glove-25d": "ade20k_150_classes_glove_twitter_27b_25d_embeddings.npy",
    "glove-25d-inst-areas": "ade20k_150_classes_glove_twitter_27b_25d_embeddings.npy",
    "glove-25d-ade20k-places": "ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy",
    "glove-25d-ade20k-places-random": "ade20k_places_classes_25d_embeddings_random.npy",
    "glove-25d-ade20k-places-human-sizes": "ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy",
    "glove-25d-ade20k-places-size_shuffled": "ade20k_places_classes_glove_twitter_27b_25d_embeddings_shuffled.npy",
    "raw": None,
}


@pytest.mark.parametrize("mode", SEM_MODES)
def test_semantics_loader_bit_exact(mode, golden_digests):
    lab, _ = sem_labels(mode)
    table = load_table(SEM_TABLE[mode]) if SEM_TABLE[mode] else None
    sizes = load_table("ade20k_classes_abs_sizes.npy") if "human-sizes" in mode else None
    raw, sem = oracle.semantics_loader(mode, lab.numpy(), table, sizes)
    assert digest(raw) == golden_digests[f"sem/{mode}/raw"]
    assert digest(sem) == golden_digests[f"sem/{mode}/out"]


@pytest.mark.parametrize("mode", INST_MODES)
def test_instance_loader_bit_exact(mode, golden_digests):
    lab, areas = inst_labels(mode)
    coco = mode == "coco"
    table = load_table("coco_81_classes_maskrcnn_ordering_glove_twitter_27b_25d_embeddings.npy" if coco else
                       "ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy")
    sizes = None
    if "human_sizes" in mode:
        sizes = load_table("ade20k_classes_abs_sizes_shuffled.npy" if "shuffled" in mode else "ade20k_classes_abs_sizes.npy")
    raw, emb, ar = oracle.instance_loader(mode, lab.numpy(), areas.numpy(), table, 0 if coco else 100, sizes)
    assert digest(raw) == golden_digests[f"inst/{mode}/raw"]
    assert digest(emb) == golden_digests[f"inst/{mode}/emb"]
    assert digest(ar) == golden_digests[f"inst/{mode}/areas"]


def test_head_matches_reference(golden):
    m = make_model(insertion_point="input", semantics_mode=None, instance_segmentation_mode=None)
    sd = {k: v for k, v in m.state_dict().items()}
    x = synthetic.decoder_features(2, 128, 176, 192, seed=21)
    with torch.no_grad():
        tgt = oracle.patch_transformer(x, sd)
        widths, ram = oracle.mvit(x, sd)
        edges, pred = oracle.head(x, sd, 1e-3, 10)
    np.testing.assert_allclose(tgt.numpy(), golden["head/tgt"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(widths.numpy(), golden["head/widths"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(edges.numpy(), golden["head/edges"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ram[:, :, ::16, ::16].numpy(), golden["head/ram_sub"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(pred.numpy(), golden["head/pred"], rtol=1e-4, atol=1e-5)


def test_losses_match_reference(golden):
    for name, (b, h, w) in {"a": (2, 104, 136), "b": (3, 64, 96)}.items():
        depth = synthetic.depth(b, h, w, seed=52)
        pred = torch.from_numpy(golden[f"loss/{name}/pred"])
        edges = torch.from_numpy(golden[f"loss/{name}/edges"])
        s = oracle.silog(pred, depth, mask=depth > 1e-3, interpolate=True)
        np.testing.assert_allclose(s.numpy(), golden[f"loss/{name}/silog"], rtol=1e-6)
        s2 = oracle.silog(torch.nn.functional.interpolate(pred, depth.shape[-2:], mode="nearest"), depth.clamp_min(0.2),
                          mask=None, interpolate=False)
        np.testing.assert_allclose(s2.numpy(), golden[f"loss/{name}/silog_nomask_noint"], rtol=1e-6)
        c = oracle.bins_chamfer(edges, depth)
        np.testing.assert_allclose(c.numpy(), golden[f"loss/{name}/chamfer"], rtol=1e-6)


def test_insertion_matches_reference(golden):
    from helpers import load_table as T
    cases = {
        "cfg2": dict(semantics_mode="glove-25d-ade20k-places", instance_segmentation_mode=None),
        "cfg3": dict(semantics_mode="glove-25d", instance_segmentation_mode="ade20k_swin_human_sizes"),
        "areas": dict(semantics_mode="glove-25d-inst-areas", instance_segmentation_mode="coco"),
        "hsizes": dict(semantics_mode="glove-25d-ade20k-places-human-sizes", instance_segmentation_mode="ade20k_swin"),
    }
    for name, kw in cases.items():
        m = make_model(insertion_point="input", image="rgb", **kw)
        sd = m.state_dict()
        h = w = 32
        img = synthetic.image(2, h, w, seed=41)
        smode, imode = kw["semantics_mode"], kw["instance_segmentation_mode"]
        slab, _ = sem_labels(smode, 2, h, w, seed=42, n_rect=(5, 10))
        sizes = T("ade20k_classes_abs_sizes.npy")
        _, sem = oracle.semantics_loader(smode, slab.numpy(), T(SEM_TABLE[smode]), sizes if "human-sizes" in smode else None)
        args = dict(semantics=torch.from_numpy(sem))
        if imode:
            ilab, iar = inst_labels(imode, 2, h, w, seed=43, n_rect=(5, 10))
            coco = imode == "coco"
            tab = T("coco_81_classes_maskrcnn_ordering_glove_twitter_27b_25d_embeddings.npy" if coco else
                    "ade20k_places_classes_glove_twitter_27b_25d_embeddings.npy")
            _, emb, ar = oracle.instance_loader(imode, ilab.numpy(), iar.numpy(), tab, 0 if coco else 100,
                                                sizes if "human_sizes" in imode else None)
            args.update(instance_labels=torch.from_numpy(emb), instance_areas=torch.from_numpy(ar))
        with torch.no_grad():
            x = oracle.input_insertion(sd, img, smode, imode, "rgb", **args)
        np.testing.assert_allclose(x.numpy(), golden[f"insert/{name}"], rtol=1e-6, atol=1e-7)


def test_chamfer_edge_cases(golden):
    depth = synthetic.depth(2, 64, 96, seed=53, all_valid=True)
    depth[1, :, 1:, :] = 0.0
    depth[1, :, 0, 5:] = 0.0
    pred = torch.from_numpy(golden["loss/edge/pred"])
    edges = torch.linspace(1e-3, 10, 257).repeat(2, 1).contiguous()
    np.testing.assert_allclose(oracle.silog(pred, depth, mask=depth > 1e-3).numpy(), golden["loss/edge/silog"], rtol=1e-6)
    np.testing.assert_allclose(oracle.bins_chamfer(edges, depth).numpy(), golden["loss/edge/chamfer"], rtol=1e-6)
    # an image with no valid target gives NaN (0/0), as the reference's mean over an empty cloud
    depth[0] = 0.0
    assert torch.isnan(oracle.bins_chamfer(edges, depth))


def test_chamfer_oracle_vs_independent_nearest_neighbour():
    """pytorch3d (v0.6.1, environment.yml:104) is not installable here, so the restatement of its chamfer_distance is pinned from a
    second side as well: the published definition -- per cloud, mean over x of the squared distance to the nearest y plus mean
    over the VALID y of the squared distance to the nearest x, then the batch mean of each -- evaluated with scipy's cKDTree (an
    independent nearest-neighbour search) on ragged, zero-padded clouds in float64."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(77)
    n, p1, p2 = 3, 256, 5000
    x = np.sort(rng.uniform(1e-3, 10.0, size=(n, p1, 1)), axis=1)
    y = rng.uniform(1e-3, 10.0, size=(n, p2, 1))
    lengths = np.array([p2, 1234, 17])
    for i, li in enumerate(lengths):
        y[i, li:] = 0.0   # pad_sequence's zero padding: must never match and never count
    got, _ = oracle.chamfer_distance(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(lengths))
    cx = cy = 0.0
    for i, li in enumerate(lengths):
        yi = y[i, :li]
        cx += float((cKDTree(yi).query(x[i])[0] ** 2).mean())
        cy += float((cKDTree(x[i]).query(yi)[0] ** 2).mean())
    want = cx / n + cy / n
    assert abs(float(got) - want) <= 1e-12 * abs(want)
    # and the BinsChamferLoss wrapper (loss.py:33-46): centres from edges, targets >= 1e-3 only
    edges = torch.from_numpy(np.concatenate([np.full((n, 1), 1e-3), np.sort(rng.uniform(1e-3, 10.0, size=(n, 256)), axis=1)], 1))
    depth = torch.from_numpy(rng.uniform(0.0, 10.0, size=(n, 1, 24, 32)) * (rng.uniform(size=(n, 1, 24, 32)) > 0.2))
    got = float(oracle.bins_chamfer(edges, depth))
    c = 0.5 * (edges[:, 1:] + edges[:, :-1]).numpy()
    want = 0.0
    for i in range(n):
        t = depth[i].flatten().numpy()
        t = t[t >= 1e-3][:, None]
        want += float((cKDTree(t).query(c[i][:, None])[0] ** 2).mean()) / n + float((cKDTree(c[i][:, None]).query(t)[0] ** 2).mean()) / n
    assert abs(got - want) <= 1e-12 * abs(want)


def test_eval_metrics_oracle_vs_reference_golden():
    """(f)2: oracle.compute_errors (float64 restatement) and oracle.eval_epilogue vs the reference's own
    utils.compute_errors outputs recorded by tests/golden/make_golden_eval.py."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden_eval as mg
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "golden_eval.npz"))
    for name, (b, h, w, H, W, lo, hi, garg, eigen, ds) in mg.CASES.items():
        pred, gt = mg.case_inputs(name)
        for i in range(b):
            g, p = oracle.eval_epilogue(pred[i:i + 1], gt[i:i + 1], lo, hi, garg, eigen, ds)
            m = oracle.compute_errors(g, p)
            row = np.array([m[k] for k in mg.KEYS] + [g.size], dtype=np.float64)
            np.testing.assert_allclose(row, gold[name][i], rtol=2e-5, atol=2e-5)
