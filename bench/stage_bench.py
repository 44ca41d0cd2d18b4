"""libshipyard_stage throughput: file -> pinned arena -> HBM with 1 / 2 / 4 / 8 workers, pinned -> HBM tickets, against the plain
pinned cudaMemcpyAsync (torch copy_) the PCIe link allows.  Files live in tmpfs (/dev/shm) so the page cache is the source; the
numbers are the H2D path's, not a disk's (bench/diskbench measures the file system)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from batch_shipyard_b200.ops.stage import Stager


def main():
    total = int(os.environ.get("STAGE_BENCH_BYTES", str(2 << 30)))
    nfiles = 8
    d = "/dev/shm/shipyard_stage_bench"
    os.makedirs(d, exist_ok=True)
    blob = os.urandom(1 << 20)
    paths = []
    for i in range(nfiles):
        p = os.path.join(d, f"f{i}.bin")
        if not os.path.exists(p) or os.path.getsize(p) != total // nfiles:
            with open(p, "wb") as f:
                for _ in range(total // nfiles >> 20):
                    f.write(blob)
        paths.append(p)
    dst = torch.empty(total, dtype=torch.uint8, device="cuda")
    per = total // nfiles
    row = {"bytes": total, "files": nfiles}
    # reference: one big pinned buffer, torch's cudaMemcpyAsync
    pin = torch.empty(total, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        dst.copy_(pin, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter(); dst.copy_(pin, non_blocking=True); torch.cuda.synchronize(); t1 = time.perf_counter()
    row["pinned_memcpy_gbs"] = round(total / (t1 - t0) / 1e9, 1)
    for workers in (1, 2, 4, 8):
        st = Stager(0, arena_bytes=workers * 2 * (16 << 20), concurrency=workers)
        for rep in range(2):
            t0 = time.perf_counter()
            ts = [st.submit_file(p, dptr=dst.data_ptr() + i * per) for i, p in enumerate(paths)]
            for t in ts:
                st.wait(t)
            t1 = time.perf_counter()
            for t in ts:
                st.release(t)
        row[f"file_to_hbm_w{workers}_gbs"] = round(total / (t1 - t0) / 1e9, 1)
        st.close()
    st = Stager(0, arena_bytes=8 << 20, concurrency=2)
    for rep in range(2):
        t0 = time.perf_counter()
        ts = [st.submit_pinned(pin.data_ptr() + i * per, per, dst.data_ptr() + i * per) for i in range(nfiles)]
        for t in ts:
            st.wait(t)
        t1 = time.perf_counter()
        for t in ts:
            st.release(t)
    row["pinned_tickets_gbs"] = round(total / (t1 - t0) / 1e9, 1)
    st.close()
    print(json.dumps(row), flush=True)
    for p in paths:
        os.remove(p)


if __name__ == "__main__":
    main()
