"""Host glue (csrc/hostglue.cpp): the in-place shrink of the per-image views that ends `non_max_suppression`
(the reference's per-image list with None for empty images, model/YOLOV3.py:306,333).  Host-only code: runs without a GPU."""
import pytest
import torch

from objectdetectionpl_b200 import _lib as L


def test_finish_views_shrinks_in_place_and_maps_empty_images_to_none():
    H = L.hostglue()
    B, n_pad = 6, 40
    rows = torch.arange(B * n_pad * 7, dtype=torch.float32).view(B, n_pad, 7)
    counts = torch.tensor([3, 0, 40, 1, 0, 17], dtype=torch.int32)
    views = list(rows.unbind(0))
    out = H.finish_views(views, counts.data_ptr(), 7)
    assert isinstance(out, list) and len(out) == B
    for b, k in enumerate(counts.tolist()):
        if k == 0:
            assert out[b] is None
            continue
        assert out[b] is views[b]                                  # the very same tensor objects, shrunk
        assert out[b].shape == (k, 7) and out[b].is_contiguous() and out[b].data_ptr() == rows[b].data_ptr()
        assert torch.equal(out[b], rows[b, :k])
    kept = H.finish_views(list(rows.unbind(0)), counts.data_ptr(), 7, True)   # keep_empty: [0, 7] tensors instead of None
    assert kept[1].shape == (0, 7) and kept[4].shape == (0, 7) and torch.equal(kept[3], rows[3, :1])
    idx = torch.arange(B * n_pad, dtype=torch.int64).view(B, n_pad)
    out = H.finish_views(list(idx.unbind(0)), counts.data_ptr(), 0)
    assert out[1] is None and out[4] is None and torch.equal(out[5], idx[5, :17]) and out[2].shape == (40,)


def test_finish_views_rejects_counts_beyond_the_padded_view_and_non_tensors():
    H = L.hostglue()
    rows = torch.zeros(2, 8, 7)
    counts = torch.tensor([9, 1], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="does not fit"):
        H.finish_views(list(rows.unbind(0)), counts.data_ptr(), 7)
    with pytest.raises(TypeError):
        H.finish_views([1, 2], torch.tensor([1, 1], dtype=torch.int32).data_ptr(), 7)
