"""collate + pinned upload: host logic on CPU, the double-buffered upload on GPU."""
import pytest
import torch

from oracle import spectral_oracle as oref
from speech_enhancement_pytorch_b200 import feeder


def make_batch(lengths, nch=2, nspk=1, seed=0):
    g = torch.Generator().manual_seed(seed)
    return [(torch.randn(nch, n, generator=g), torch.randn(nspk, nch, n, generator=g), {}, {}, f"utt{i}")
            for i, n in enumerate(lengths)]


@pytest.mark.parametrize("drop_last", [True, False])
@pytest.mark.parametrize("lengths", [[4000, 9000, 2500], [8000], [1000, 8001, 16000]])
def test_collate_pad_matches_reference(lengths, drop_last):
    batch = make_batch(lengths)
    mix, src, nsegs = feeder.collate_pad(batch, 4000, drop_last)
    mix_ref, src_ref, idx_ref = oref.collate_fn_pad_ref(batch, 4000, drop_last)
    assert nsegs == idx_ref
    assert torch.equal(mix, mix_ref) and torch.equal(src, src_ref)
    # writing into (oversized, dirty) staging buffers gives the same bytes
    out = (torch.full((32, 2, 4000), 7.0), torch.full((32, 1, 2, 4000), 7.0))
    mix2, src2, _ = feeder.collate_pad(batch, 4000, drop_last, out=out)
    assert torch.equal(mix2, mix_ref) and torch.equal(src2, src_ref)


@pytest.mark.gpu
def test_pinned_feeder_uploads_every_batch_unchanged():
    loader = [make_batch([4000, 9000, 2500], seed=s) for s in range(5)]
    f = feeder.PinnedFeeder(loader, 4000, "cuda", max_segments=16, channels=2, speakers=1)
    seen = 0
    for (mix, src, nsegs), batch in zip(f, loader):
        mix_ref, src_ref, idx_ref = oref.collate_fn_pad_ref(batch, 4000, True)
        assert mix.is_cuda and nsegs == idx_ref
        assert torch.equal(mix.cpu(), mix_ref) and torch.equal(src.cpu(), src_ref)
        seen += 1
    assert seen == 5 and f.bytes_per_batch > 0


@pytest.mark.parametrize("tag,drop", [("drop", True), ("pad", False)])
def test_collate_pad_matches_reference_run_golden(tag, drop):
    """feeder.collate_pad against the output of the REAL reference's collate_fn_pad (src/distrib.py:38-98) on the
    same clips: tests/golden/collate.npz, made by tests/golden/make_golden.py."""
    import numpy as np
    from conftest import golden
    gd = golden("collate")
    batch = [(torch.from_numpy(gd[f"mix_{i}"]), torch.from_numpy(gd[f"src_{i}"])) for i in range(4)]
    mix, src, nsegs = feeder.collate_pad(batch, int(gd["segment_length"]), drop)
    assert np.array_equal(mix.numpy(), gd[f"batch_mix_{tag}"])
    assert np.array_equal(src.numpy(), gd[f"batch_src_{tag}"])
    assert list(nsegs) == list(gd[f"index_{tag}"])


def test_collate_pad_rejects_batches_larger_than_the_staging_buffers():
    batch = make_batch([9000, 9000])
    out = (torch.empty(3, 2, 4000), torch.empty(3, 1, 2, 4000))
    with pytest.raises(ValueError, match="max_segments"):
        feeder.collate_pad(batch, 4000, True, out=out)
