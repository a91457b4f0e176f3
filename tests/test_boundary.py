"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
that include/dfvit.h declares, the topology / blob-layout queries (pure host code) agree with
the oracle's module tree, the product never imports the oracle, and the nn.Module drop-in has
the reference's state_dict layout.  No kernels are launched here."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def d():
    import __graft_entry__ as g
    if not os.path.isfile(os.path.join(ROOT, "deepfake_vit_b200", "libdfvit.so")):
        g.build()
    import deepfake_vit_b200
    return deepfake_vit_b200


def test_library_exports_every_declared_symbol(d):
    header = open(os.path.join(ROOT, "include", "dfvit.h")).read()
    declared = set(re.findall(r"\b(dfv_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    raw = ctypes.CDLL(d._lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in dfvit.h but not exported"
    assert declared == set(d._lib.EXPORTS), declared ^ set(d._lib.EXPORTS)
    assert d._lib.lib.dfv_version() >= 100


def test_topology_matches_oracle(d):
    from oracle import calibrate, refmodel
    om = calibrate.build(refmodel.get_oracle(), "default")
    blocks = om.feature_extractor.backbone.backbone._blocks
    infos = d._lib.b4_blocks()
    assert len(infos) == len(blocks) == 32
    for info, blk in zip(infos, blocks):
        a = blk._block_args
        dw = blk._depthwise_conv
        assert (info.c_in, info.c_out, info.kernel, info.stride) == (a.input_filters, a.output_filters, a.kernel_size, a.stride)
        assert info.c_mid == dw.in_channels and info.se_squeeze == blk._se_reduce.out_channels
        pad = dw.static_padding.padding
        assert (info.pad_lo, info.pad_hi) == (pad[0], pad[1]) == (pad[2], pad[3])
        assert bool(info.has_expand) == hasattr(blk, "_expand_conv")
        assert bool(info.has_skip) == (a.stride == 1 and a.input_filters == a.output_filters)
    ho, wo = ctypes.c_int(), ctypes.c_int()
    for size, want in ((380, 12), (224, 7), (160, 5)):
        assert d._lib.lib.dfv_b4_output_hw(size, size, ctypes.byref(ho), ctypes.byref(wo)) == 0
        assert (ho.value, wo.value) == (want, want)
        with torch.no_grad():
            assert om.feature_extractor.backbone.get_feature_maps(torch.zeros(1, 3, size, size)).shape[-1] == want


def test_blob_layout_is_consistent(d):
    for dtype in (0, 1):
        total = d._lib.lib.dfv_blob_bytes(dtype)
        spans = []
        for i, info in enumerate(d._lib.b4_blocks()):
            for kind in range(d._lib.W_EXPAND, d._lib.W_PROJECT_BIAS + 1):
                if kind in (d._lib.W_EXPAND, d._lib.W_EXPAND_BIAS) and not info.has_expand:
                    with pytest.raises(d._lib.DfvError):
                        d._lib.blob_slot(dtype, i, kind)
                    continue
                off, n = d._lib.blob_slot(dtype, i, kind)
                es = (2 if dtype else 4) if kind in (d._lib.W_EXPAND, d._lib.W_PROJECT) else 4
                spans.append((off, off + n * es))
        for kind in (d._lib.W_STEM, d._lib.W_STEM_BIAS, d._lib.W_HEAD, d._lib.W_HEAD_BIAS):
            off, n = d._lib.blob_slot(dtype, -1, kind)
            spans.append((off, off + n * ((2 if dtype else 4) if kind == d._lib.W_HEAD else 4)))
        spans.sort()
        assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] <= total
        assert all(s[0] % 256 == 0 for s in spans)
    assert d._lib.lib.dfv_infer_workspace_bytes(1, 256, 380, 380) < 8 << 30


def test_state_dict_layout_and_seeded_init_equal_the_reference(d):
    from oracle import calibrate, refmodel
    om = calibrate.build(refmodel.get_oracle(), "default")
    torch.manual_seed(calibrate.SEED_MODEL)
    m = d.DeepfakeDetectionModel(**refmodel.MODEL_CONFIG)
    sa, sb = om.state_dict(), m.state_dict()
    assert list(sa) == list(sb) and len(sb) == 731
    assert all(sa[k].shape == sb[k].shape and torch.equal(sa[k], sb[k]) for k in sa)
    m.load_state_dict(sa, strict=True)
    assert sum(p.numel() for p in m.parameters()) == 18_939_345
    assert m.feature_extractor.feature_dim == 1792 and m.num_classes == 2
    # constructor contract: config dict wins over top-level `pretrained`; attention can be switched off
    m2 = d.DeepfakeDetectionModel(feature_extractor_config={"pretrained": False, "use_attention": False})
    assert m2.feature_extractor.attention is None
    m3 = d.DeepfakeDetectionModel(feature_extractor_config={"attention_config": {"use_landmark": True, "use_spatial": False, "use_channel": True}},
                                  classifier_hidden_dims=[64], num_classes=3)
    assert not hasattr(m3.feature_extractor.attention, "spatial_attn") and m3.classifier[-1].out_features == 3


def test_no_cpu_fallback_and_no_oracle_in_product(d):
    m = d.DeepfakeDetectionModel(feature_extractor_config={"pretrained": False}).eval()
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 3, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU path"):
        d.CombinedLoss({"ce": 1.0})(torch.zeros(2, 2), torch.zeros(2, dtype=torch.int64))
    if not torch.cuda.is_available():
        assert d._lib.lib.dfv_device_check() != 0          # launch entries refuse to run without sm_100
    pkg = os.path.join(ROOT, "deepfake_vit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "efficientnet_pytorch" not in src or f == "model.py" and "import efficientnet_pytorch" not in src


def test_tile_planners_stay_within_the_sm(d):
    """Host-side planners (no launch): every 1x1-conv GEMM and depthwise layer of B4, at the benchmark batches and a few
    odd ones, gets a plan that fits one SM -- <= 227 KB of shared memory, >= 2 pipeline stages, a grid of at most a
    few CTAs per SM -- and the weight-stationary GEMM plan appears only where its weight tile fits."""
    lib = d._lib.lib
    out = (ctypes.c_int * 10)()
    dw = (ctypes.c_int * 10)()
    for B in (1, 3, 64, 256):
        size = 190
        for b in d._lib.b4_blocks():
            hw_in = size * size
            size_out = (size + b.stride - 1) // b.stride
            hw_out = size_out * size_out
            shapes = []
            if b.has_expand:
                shapes.append((B * hw_in, b.c_in, b.c_mid, 0))
            shapes.append((B * hw_out, b.c_mid, b.c_out, 1))
            for (M, K, N, gated) in shapes:
                assert lib.dfv_gemm_plan_info(ctypes.c_longlong(M), K, N, gated, out) == 0, (B, M, K, N)
                bn, res, stages, nbuf, grid, tpc, smem, ntn, cl, nsub = list(out)
                assert 16 <= bn <= 256 and bn % 16 == 0 and ntn * bn >= N
                assert stages >= 2 and nbuf in (1, 2) and 0 < smem <= 227 * 1024
                assert 1 <= grid <= 148 and tpc >= 1 and grid * tpc * (1 if res else 1) >= 1
                assert nsub in (1, 2) and (nsub == 1 or (ntn == 2 and not res and gated and cl == 2 and 2 * bn <= 512 and tpc % 2 == 0 and stages >= 3))
                assert cl in (1, 2, 4) and grid % cl == 0 and (bn // cl) % 8 == 0 and not (res and cl > 1)
                if res:     # the resident weight tile leaves room for the activation ring (gated: one N tile, <= 96 KB)
                    assert ((K + 63) // 64) * 64 * bn * 2 <= (96 if gated else 160) * 1024 and (not gated or ntn == 1)
            assert lib.dfv_dwconv_plan_info(1, B, size, size, b.c_mid, b.kernel, b.stride, b.pad_lo, b.pad_hi, dw) == 0
            assert 0 < dw[5] <= 227 * 1024 and dw[4] <= 256 and dw[8] >= 1 and 1 <= dw[9] <= 148 * 4
            size = size_out
