"""Multi-GPU check of the library's own collective (tecdsa_nccl_* + tecdsa_gather_results / tecdsa_gg20_offline_records): every rank
runs a different small batch of offline sessions, gathers the 256-byte records with ONE ncclAllGather issued by the C library, and
the result is compared on every rank with torch.distributed's all_gather of the same records.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/nccl_gather_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import __graft_entry__ as entry
from tests.golden import fixtures
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = entry.load_package()
from mpecdsa_b200 import gg20
eng = pkg.Engine(local)
keysets = fixtures.load_all_keysets()[:2]
ks = gg20.KeySets(eng, keysets)
n = 96
sess, rnd = gg20.synthetic_batch(keysets, n, 1000 + rank)            # a different batch on every rank
U = 2 * n
idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    idt.copy_(torch.from_numpy(eng.nccl_unique_id()))
dist.broadcast(idt, 0)
comm = eng.nccl_comm_create(idt.cpu().numpy(), world, rank)
# path 1: the one-call entry point with host buffers (H2D -> rounds -> pack -> gather -> D2H)
h_all = np.zeros((world, U, 256), np.uint8)
eng.offline_records(ks, comm, sess, n, rnd, h_all, pkg.HOST)
# path 2: device buffers, explicit pack + gather
d_sess, d_rnd = torch.from_numpy(sess.view(np.int32)).cuda(), torch.from_numpy(rnd.view(np.int32)).cuda()
d_status = torch.empty(U, dtype=torch.uint8, device="cuda")
d_R, d_sigma = torch.empty((U, 16), dtype=torch.int32, device="cuda"), torch.empty((U, 8), dtype=torch.int32, device="cuda")
d_tvec, d_digest = torch.empty((U, 32), dtype=torch.int32, device="cuda"), torch.empty((U, 8), dtype=torch.int32, device="cuda")
d_rec = torch.empty((U, 256), dtype=torch.uint8, device="cuda")
d_all = torch.zeros((world, U, 256), dtype=torch.uint8, device="cuda")
gg20.offline_raw(eng, ks, d_sess, n, d_rnd, d_status, d_R, d_sigma, d_tvec, d_digest, pkg.DEVICE)
eng.pack_records(d_status, d_R, d_sigma, d_tvec, d_digest, d_rnd, U, d_rec)
eng.gather_results(comm, d_rec, U, d_all)
eng.sync()
# reference: torch.distributed's own all_gather of the same local records
ref = [torch.empty_like(d_rec) for _ in range(world)]
dist.all_gather(ref, d_rec)
ref = torch.stack(ref)
ok = bool((d_all == ref).all()) and bool((torch.from_numpy(h_all).cuda() == ref).all()) and not bool(d_all[:, :, 0].any())
distinct = all(not bool((ref[0] == ref[r]).all()) for r in range(1, world))
flag = torch.tensor([1 if (ok and distinct) else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("NCCL_GATHER_CHECK", {"world": world, "units_per_rank": U, "all_ranks_ok": bool(flag.item()), "records_differ_between_ranks": distinct}, flush=True)
eng.nccl_comm_destroy(comm)
ks.free(); eng.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() else 1)
