"""libshipyard_stage on a GPU: file -> pinned arena -> HBM and pinned -> HBM tickets, event chaining, against plain byte comparison."""

import pytest
import torch

pytestmark = pytest.mark.gpu

if not torch.cuda.is_available():
    pytest.skip("needs a GPU", allow_module_level=True)


@pytest.mark.parametrize("workers", [1, 4])
def test_file_to_hbm_roundtrip(tmp_path, workers):
    from batch_shipyard_b200.ops.stage import Stager
    g = torch.Generator().manual_seed(workers)
    sizes = [(5 << 20) + 123, 4096, 1, (17 << 20) + 1]
    blobs = [torch.randint(0, 256, (n,), dtype=torch.uint8, generator=g) for n in sizes]
    paths = []
    for i, b in enumerate(blobs):
        p = tmp_path / f"blob{i}.bin"
        p.write_bytes(bytes(b.numpy()))
        paths.append(str(p))
    st = Stager(0, arena_bytes=16 << 20, concurrency=workers)
    dst = [torch.zeros(n, dtype=torch.uint8, device="cuda") for n in sizes]
    tickets = [st.submit_file(p, dptr=d.data_ptr()) for p, d in zip(paths, dst)]
    own = st.submit_file(paths[0], offset=1000, nbytes=2 << 20)          # stager-owned destination, a byte range of the file
    stream = torch.cuda.current_stream()
    for t in tickets + [own]:
        st.stream_wait(t, stream.cuda_stream)                             # the consumer is event-chained, the host does not wait
    sums = [d.to(torch.int64).sum() for d in dst]                         # enqueued behind the copies on the current stream
    for d, b, s_ in zip(dst, blobs, sums):
        assert int(s_) == int(b.to(torch.int64).sum())
        assert torch.equal(d.cpu(), b)
    assert st.wait(own) and st.query(own).bytes == 2 << 20
    from batch_shipyard_b200.ops.coll import _CudaView
    view = torch.as_tensor(_CudaView(st.ptr(own), 2 << 20, st), device="cuda")
    assert torch.equal(view.cpu(), blobs[0][1000:1000 + (2 << 20)])
    stats = st.stats()
    assert stats["bytes_staged"] == sum(sizes) + (2 << 20) and stats["memcpy_calls"] >= len(sizes) + 1
    for t in tickets + [own]:
        st.release(t)
    st.close()


def test_pinned_ticket_waits_for_consumer_event():
    """The double-buffered input path: the copy into a slot starts only after the kernel that read the slot has finished."""
    from batch_shipyard_b200.ops.stage import Stager
    st = Stager(0, arena_bytes=8 << 20, concurrency=1)
    n = 32 << 20
    host = [torch.full((n,), v, dtype=torch.uint8).pin_memory() for v in (1, 2)]
    dev = torch.zeros(n, dtype=torch.uint8, device="cuda")
    cur = torch.cuda.current_stream()
    t0 = st.submit_pinned(host[0].data_ptr(), n, dev.data_ptr())
    st.stream_wait(t0, cur.cuda_stream)
    torch.cuda._sleep(50_000_000)                       # a long "step" that reads the slot afterwards
    first = dev.to(torch.int64).sum()
    consumed = torch.cuda.Event(); consumed.record(cur)
    t1 = st.submit_pinned(host[1].data_ptr(), n, dev.data_ptr(), wait_event=consumed.cuda_event)   # must not overtake `first`
    st.stream_wait(t1, cur.cuda_stream)
    second = dev.to(torch.int64).sum()
    assert int(first) == n * 1 and int(second) == n * 2
    for t in (t0, t1):
        st.wait(t); st.release(t)
    st.close()


def test_input_stager_uses_the_native_library():
    from batch_shipyard_b200.models.resnet import resnet_tiny
    from batch_shipyard_b200.ops.coll import Communicator
    from batch_shipyard_b200.parallel.ddp import FusedDataParallelTrainer
    comm = Communicator(0, 1, device=0, heap_bytes=256 << 20)
    tr = FusedDataParallelTrainer(resnet_tiny(10), comm, (8, 3, 64, 64), 10, lr=0.05, use_graph=False)
    stg = tr.make_stager()
    stg.fill_synthetic(0)
    for hy in stg.host_y:
        hy.remainder_(10)
    losses = []
    stg.prefetch(0)
    for i in range(4):
        stg.run_step(i % 2)
        stg.prefetch((i + 1) % 2)
        losses.append(stg.read_loss(i % 2))
    assert all(l == l for l in losses)
    assert torch.equal(stg.dev_x[1].cpu(), stg.host_x[1])
    stg.close()
    summ = stg.staging_summary_cached
    assert "libshipyard_stage" in summ["path"] and summ["memcpy_calls"] >= 8
    maps = open("/proc/self/maps").read()
    assert "libshipyard_stage.so" in maps
    comm.close()
