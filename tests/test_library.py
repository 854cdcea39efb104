"""CPU checks: the shared object builds/loads and exports every symbol the header declares; the drop-in
modules keep the reference's state-dict layout; the product path refuses CPU tensors (no fallback)."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

from murcl_b200 import _lib, synth

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    hdr = (ROOT / "include" / "murcl_b200.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(murcl_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_header_symbols():
    from murcl_b200 import build
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/murcl_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header drifted"
    loaded = _lib.load()
    assert loaded.murcl_version() == 1
    assert loaded.murcl_launch_count() >= 0


def test_argument_validation_without_gpu():
    """Bad arguments are rejected before any CUDA call, with a message."""
    lib = _lib.load()
    rc = lib.murcl_pack_gather(None, 0, 8, None, 1, 4, None, None, None, 0, None)
    assert rc == -1 and b"null" in lib.murcl_last_error()
    rc = lib.murcl_linear_fwd(1, 1, None, 1, 4, 0, 4, 0, 0, 0, 0, None, None)
    assert rc == -1
    rc = lib.murcl_ntxent_fwd_bwd(1, 0, 4, 1.0, 1, None, None, 1, None)
    assert rc == -1
    with pytest.raises(_lib.MurclError):
        _lib.check(rc, "murcl_ntxent_fwd_bwd")


def test_state_dict_layout_matches_reference_names():
    from murcl_b200.dropin import abmil, clam, dsmil, rlmil
    abmil.ABMIL(24, L=32, D=16).load_state_dict(synth.abmil_state(24, 32, 16, 2), strict=True)
    for gate in (True, False):
        for dropout in (True, False):
            m = clam.CLAM_SB(gate=gate, dropout=dropout, n_classes=3, in_dim=24)
            m.load_state_dict(synth.clam_state(24, "small", gate, dropout, 3), strict=True)
    clam.CLAM_SB(size_arg="big", in_dim=24).load_state_dict(synth.clam_state(24, "big"), strict=True)
    net = dsmil.MILNet(dsmil.FCLayer(40, 2), dsmil.BClassifier(40, 2))
    net.load_state_dict(synth.dsmil_state(40, 2), strict=True)
    rlmil.Full_layer(24, 40, True, 12).load_state_dict(synth.full_layer_state(24, 40, 12), strict=True)
    ac = rlmil.ActorCritic(32, 32, 24, False, 0.5, 6)
    ac.load_state_dict(synth.actor_state(32, 24, 6), strict=True)


def test_no_cpu_fallback():
    from murcl_b200 import ops
    from murcl_b200.dropin import abmil, clam, datasets, losses
    with pytest.raises(_lib.MurclError):
        abmil.ABMIL(16, L=16, D=8)(torch.randn(2, 5, 16))
    with pytest.raises(_lib.MurclError):
        clam.CLAM_SB(in_dim=16)(torch.randn(1, 20, 16))
    with pytest.raises(_lib.MurclError):
        losses.NT_Xent(2, 1.0)(torch.randn(2, 4), torch.randn(2, 4))
    with pytest.raises(_lib.MurclError):
        ops.linear(torch.randn(2, 4), torch.randn(3, 4))
    with pytest.raises(_lib.MurclError):
        datasets.get_feats([torch.randn(1, 9, 4)], [[[0, 1, 2], [3, 4, 5, 6, 7, 8]]], torch.rand(1, 2), 4)


def test_product_does_not_import_oracle():
    for path in (ROOT / "murcl_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path


def test_clam_instance_plan_layout():
    """Host-side group layout of the instance loss (clam.py:146-166): label -> targets."""
    from murcl_b200.dropin import clam

    class R:
        B = 2
        sizes = [20, 30]

        class rows:
            device = "cpu"

    m = clam.CLAM_SB(n_classes=2, subtyping=True, in_dim=16)
    plan = m._instance_plan(R, [1, 0])
    assert plan["targets"].tolist() == [0] * 8 + [1] * 8 + [0] * 8 + [1] * 8 + [0] * 8 + [0] * 8
    assert plan["group_cls"].tolist() == [0, 1, 0, 1]
    assert plan["groups"] == [(0, 0, False), (0, 1, True), (1, 0, True), (1, 1, False)]
    m2 = clam.CLAM_SB(n_classes=2, subtyping=False, in_dim=16)
    assert m2._instance_plan(R, [1, 0])["group_cls"].tolist() == [1, 0]
    R.sizes = [5, 30]
    with pytest.raises(RuntimeError):
        m._instance_plan(R, [1, 0])


def test_attnpool_supported_shapes():
    """Host-side shape gate of the fused attention-pooling kernel (no device work)."""
    from murcl_b200 import _lib
    lib = _lib.load()
    assert lib.murcl_attnpool_supported(512, 128, 0, _lib.BF16) == 1      # ABMIL
    assert lib.murcl_attnpool_supported(512, 256, 1, _lib.BF16) == 1      # CLAM_SB small, gated
    assert lib.murcl_attnpool_supported(512, 384, 1, _lib.BF16) == 0      # CLAM_SB big, gated: 768 columns
    assert lib.murcl_attnpool_supported(1024, 128, 0, _lib.BF16) == 0
    assert lib.murcl_attnpool_supported(512, 128, 0, _lib.F32) == 0
    assert lib.murcl_attnpool_workspace(1000, 3, 512) == (8 + 3 + 1) * 516


def test_attnpool_record_index_is_unique_per_tile_bag_incidence():
    """The fused pooling kernel writes one partial record per (128-row tile, bag) incidence at index tile + bag and the
    merge kernel reads records tile + bag for tile in [off[b] // 128, (off[b+1] - 1) // 128] (csrc/attnpool.cu).  Both
    indices only grow along the rows, so the sum is unique and stays inside the (tiles + B + 1)-record workspace - also
    with empty bags, bags smaller than a tile and boundaries that coincide with tile boundaries."""
    import random
    rng = random.Random(7)
    for _ in range(200):
        B = rng.randint(1, 12)
        sizes = [rng.choice([0, 1, 5, 127, 128, 129, 256, 300, rng.randint(0, 700)]) for _ in range(B)]
        off = [0]
        for n in sizes:
            off.append(off[-1] + n)
        n_rows = off[-1]
        tiles = (n_rows + 127) // 128
        written = {}
        for t in range(tiles):                                  # kernel side: segments of tile t
            row0, last = t * 128, min(t * 128 + 127, n_rows - 1)
            seg = lambda r: max(b for b in range(B) if off[b] <= r and sizes[b] > 0 and r < off[b + 1])
            for b in range(seg(row0), seg(last) + 1):
                r_begin, r_end = max(off[b] - row0, 0), min(off[b + 1] - row0, 128)
                if r_end > r_begin:
                    assert (t + b) not in written, "two incidences share a record"
                    written[t + b] = (t, b)
        for b in range(B):                                      # merge side
            if sizes[b] == 0:
                continue
            for t in range(off[b] // 128, (off[b + 1] - 1) // 128 + 1):
                assert written.get(t + b) == (t, b), "merge reads a record of another incidence"
        assert all(k < tiles + B + 1 for k in written)


def test_host_bags_from_reference_file_formats(tmp_path):
    """.npz features + .npz labels / .json inverted lists -> pinned CSR staging.  The ranks must
    reproduce the positions get_feats slices (utils/datasets.py:293-296), also for a list that is not in ascending order."""
    import json
    import numpy as np
    from murcl_b200 import synth
    from murcl_b200.csr import HostBags, load_slide
    from oracle import murcl_oracle as O
    K, D = 4, 8
    feats, clusters, labels = synth.make_bags([50, 7, 33], D, K, seed=12)
    clusters[2][1] = clusters[2][1][::-1]                       # one list deliberately reversed (JSON slide only)
    ffiles, cfiles = [], []
    for b, (f, c, l) in enumerate(zip(feats, clusters, labels)):
        ff = tmp_path / f"slide{b}.npz"
        np.savez(ff, img_features=f.numpy())
        if b == 1:                                              # label file as written by features_clustering.py:10-16
            cf = tmp_path / f"slide{b}_clusters.npz"
            np.savez(cf, features_cluster_indices=l.numpy().reshape(-1, 1))
        else:                                                   # inverted lists as written by features_clustering.py:19-25
            cf = tmp_path / f"slide{b}.json"
            cf.write_text(json.dumps(c))
        ffiles.append(str(ff)); cfiles.append(str(cf))
    host = HostBags.from_files(ffiles, cfiles, K, pin=False)
    assert host.offsets == [0, 50, 57, 90]
    assert torch.equal(host.feats, torch.cat(feats))
    assert torch.equal(host.patch_cluster, torch.cat(labels))
    for b in range(3):
        lo, hi = host.offsets[b], host.offsets[b + 1]
        for j in range(K):
            ids = clusters[b][j]
            assert host.cluster_sizes[b, j] == len(ids)
            # rank = position inside the list the reference slices
            assert host.patch_rank[lo:hi][torch.tensor(ids, dtype=torch.long)].tolist() == list(range(len(ids)))
    # the window selection driven by (cluster, rank) equals the oracle's slicing of the lists
    act = torch.rand(1, K, generator=synth.gen(5))
    for b in range(3):
        lo, hi = host.offsets[b], host.offsets[b + 1]
        want = O.select_indices(clusters[b], hi - lo, act[0].numpy(), 16)
        starts, stops = O.select_windows([len(c) for c in clusters[b]], hi - lo, act[0].numpy(), 16)
        pc, pr = host.patch_cluster[lo:hi].long(), host.patch_rank[lo:hi].long()
        keep = (pr >= torch.as_tensor(starts)[pc]) & (pr < torch.as_tensor(stops)[pc])
        assert torch.nonzero(keep).flatten().tolist()[:16] == [int(i) for i in want]
    with pytest.raises(ValueError):
        load_slide(ffiles[0], cfiles[1], K)                     # 7 labels for 50 patches
