"""Chunked ingest (xgcm_b200/ingest.py): file order, contents, error delivery; host-only ring on CPU, the full
disk -> pinned -> device pipeline on the GPU box."""

import numpy as np
import pytest
import torch

from xgcm_b200 import ingest


def _chunks(n, shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.random(shape).astype(dtype) for _ in range(n)]


@pytest.mark.parametrize("raw", [False, True])
@pytest.mark.parametrize("depth", [2, 3])
def test_chunkstream_host_ring(tmp_path, raw, depth):
    arrays = _chunks(7, (3, 5, 8), np.float32)
    paths = ingest.write_chunks(str(tmp_path), arrays, raw=raw)
    kw = dict(shape=(3, 5, 8), dtype=np.float32) if raw else {}
    stream = ingest.ChunkStream(paths, device="cpu", depth=depth, **kw)
    seen = []
    for k, t in stream:
        np.testing.assert_array_equal(t.numpy(), arrays[k])
        seen.append(k)
    assert seen == list(range(7)) and len(stream) == 7
    assert stream.bytes_read == 7 * arrays[0].nbytes


def test_chunkstream_errors(tmp_path):
    arrays = _chunks(2, (4, 4), np.float64)
    paths = ingest.write_chunks(str(tmp_path), arrays)
    bad = ingest.write_chunks(str(tmp_path / "bad"), [np.zeros((4, 5))])
    with pytest.raises(ValueError, match="expected"):
        for _ in ingest.ChunkStream(paths + bad, device="cpu"):
            pass
    with pytest.raises(ValueError, match="explicit shape"):
        ingest.ChunkStream([str(tmp_path / "x.bin")], device="cpu")
    with pytest.raises(ValueError):
        ingest.ChunkStream([], device="cpu")
    # early exit of the consumer must not leave the reader blocked
    it = iter(ingest.ChunkStream(paths, device="cpu", depth=2))
    next(it)
    it.close()


@pytest.mark.gpu
def test_chunkstream_device_pipeline(tmp_path):
    """disk -> pinned ring -> device: every chunk arrives intact and in order while earlier chunks are being used."""
    from oracle import stencil as oracle
    from xgcm_b200 import ops

    arrays = _chunks(9, (6, 64, 256), np.float32, seed=3)
    paths = ingest.write_chunks(str(tmp_path), arrays)
    outs = []
    for k, t in ingest.ChunkStream(paths, depth=3):
        assert t.is_cuda
        outs.append(ops.stencil2(t, 2, "diff", 1, 0, "periodic"))  # consumer work queued behind the upload
    torch.cuda.synchronize()
    for k, o in enumerate(outs):
        np.testing.assert_array_equal(o.cpu().numpy(), oracle.stencil2("diff", arrays[k], 2, 1, 0, "periodic", 0.0))
