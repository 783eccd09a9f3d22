"""cplb_create_sharded: one process, one set of host buffers, several device pipelines (instances shard by
index, contiguous ranges, no collective).  On a one-GPU box the device list names GPU 0 several times -- the code path
(range cutting, interleaved enqueue over the pipelines, tickets spanning pipelines, one packing thread per shard for
pageable buffers) is the same; with two or more GPUs the same tests run over distinct devices."""
import ctypes as C

import numpy as np
import pytest

import centroidalplanner_b200 as cpl
from centroidalplanner_b200 import _cabi, synthetic

from helpers import assert_parity, configure, make_pair, same_bits, to_instance_major, CASES

pytestmark = pytest.mark.gpu
LAYOUTS = [cpl.INSTANCE_MAJOR, cpl.COMPONENT_MAJOR]


def device_lists():
    import torch

    n = torch.cuda.device_count()
    lists = [[0, 0, 0]]
    if n >= 2:
        lists.append(list(range(n)))
    return lists


def sharded_twin(case, devices):
    """(single-device problem, sharded problem, oracle, generator) with identical parameters."""
    prob, o, gen = make_pair(case)
    names, env_name = CASES[case][:2]
    extra = CASES[case][3] if len(CASES[case]) > 3 else None
    env = {"none": None, "ground": cpl.Ground, "superquadric": cpl.Superquadric}[env_name]
    env = env() if env is not None else None
    sh = cpl.BatchedCplProblem(names, 100.0, env, devices=devices)
    configure(sh, env, names, env_name, True, extra)
    return prob, sh, o, gen


def pinned(lib, shape):
    ptr = C.c_void_p()
    assert lib.cplb_host_alloc(int(np.prod(shape)) * 8, C.byref(ptr)) == 0
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(int(np.prod(shape)),)).reshape(shape), ptr


def test_shard_ranges_are_contiguous_and_tile_aligned(cuda_device):
    for devices in device_lists() + [[0], [0] * 8]:
        sh = cpl.BatchedCplProblem(synthetic.NAMES4, 100.0, cpl.Ground(), devices=devices)
        assert sh.GetNumShards() == len(devices)
        for N in (0, 1, 31, 32, 33, 1000, 65536, 65537, 1 << 20):
            at = 0
            for s in range(len(devices)):
                dev, b, e = sh.GetShard(s, N)
                assert dev == devices[s] and b == at and b <= e <= N
                assert b % 32 == 0 or b == N
                at = e
            assert at == N
        with pytest.raises(ValueError, match="out of range"):
            sh.GetShard(len(devices), 10)


@pytest.mark.parametrize("mode", ["pinned", "pageable", "queued"])
@pytest.mark.parametrize("layout", LAYOUTS)
@pytest.mark.parametrize("case", ["ground4", "superquadric3", "noenv8"])
def test_sharded_host_evaluation_equals_the_single_device_one(case, layout, mode, cuda_device):
    lib = _cabi.load()
    for devices in device_lists():
        prob, sh, o, gen = sharded_twin(case, devices)
        N = 100003  # ragged: the last shard is shorter, several chunks per shard
        x = gen(N)
        xin = x if layout == cpl.INSTANCE_MAJOR else np.ascontiguousarray(x.T)
        want = prob.eval(xin, g=True, jac=True, cost=True, grad=True, layout=layout)
        shp = (lambda L: (N, L)) if layout == cpl.INSTANCE_MAJOR else (lambda L: (L, N))
        keep = []
        if mode == "pageable":
            got = sh.eval(xin, g=True, jac=True, cost=True, grad=True, layout=layout)
        else:
            bufs = {}
            for key, shape in (("x", shp(o.n)), ("g", shp(o.m)), ("jac", shp(o.nnz)), ("grad", shp(o.n)), ("cost", (N,))):
                bufs[key], ptr = pinned(lib, shape)
                keep.append(ptr)
                bufs[key][...] = np.nan
            bufs["x"][...] = xin
            out = {k: bufs[k] for k in ("g", "jac", "cost", "grad")}
            if mode == "pinned":
                got = sh.eval(bufs["x"], g=True, jac=True, cost=True, grad=True, layout=layout, out=out)
            else:
                ticket, got = sh.eval_host_begin(bufs["x"], out, g=True, jac=True, cost=True, grad=True, layout=layout)
                sh.eval_host_wait(ticket)
        for key in ("g", "jac", "cost", "grad"):
            assert same_bits(got[key], want[key]), f"{case}/{devices}/{mode}: {key}"
        sub = np.arange(0, N, 97)
        assert_parity({k: to_instance_major(np.asarray(got[k]), layout)[sub] for k in got}, o.eval_batch(x[sub], nthreads=4), o,
                      f"sharded/{case}/{devices}/{mode}", x[sub])
        for ptr in keep:
            lib.cplb_host_free(ptr)


def test_sharded_queue_of_batches_and_per_instance_parameters(cuda_device):
    """Two batches in flight on a sharded problem (begin k+1, wait k), per-instance wrench arrays cut with the batch."""
    lib = _cabi.load()
    for devices in device_lists():
        prob, sh, o, gen = sharded_twin("ground4", devices)
        N = 50000
        keep, sets = [], []
        rng = np.random.default_rng(5)
        for _ in range(2):
            b = {}
            for key, shape in (("x", (N, o.n)), ("g", (N, o.m)), ("jac", (N, o.nnz)), ("wrench", (N, 6))):
                b[key], ptr = pinned(lib, shape)
                keep.append(ptr)
            sets.append(b)
        batches = [gen(N) * (1.0 + 0.01 * q) for q in range(3)]
        wr = [rng.uniform(-50, 50, (N, 6)) for _ in range(3)]
        want = [prob.eval(batches[q], g=True, jac=True, per_instance={"wrench": wr[q]}) for q in range(3)]
        got, pending = [], None
        for q in range(3):
            b = sets[q % 2]
            b["x"][...] = batches[q]
            b["wrench"][...] = wr[q]
            b["g"][...] = np.nan
            b["jac"][...] = np.nan
            ticket, _ = sh.eval_host_begin(b["x"], {"g": b["g"], "jac": b["jac"]}, g=True, jac=True, per_instance={"wrench": b["wrench"]})
            if pending is not None:
                sh.eval_host_wait(pending[0])
                got.append({k: pending[1][k].copy() for k in ("g", "jac")})
            pending = (ticket, b)
        sh.eval_host_wait(pending[0])
        got.append({k: pending[1][k].copy() for k in ("g", "jac")})
        for q in range(3):
            assert same_bits(got[q]["g"], want[q]["g"]) and same_bits(got[q]["jac"], want[q]["jac"]), (devices, q)
        for ptr in keep:
            lib.cplb_host_free(ptr)


def test_device_buffers_of_a_sharded_problem_go_through_eval_shard(cuda_device):
    import torch

    for devices in device_lists():
        prob, sh, o, gen = sharded_twin("ground8", devices)
        N = 70000
        x = gen(N)
        with pytest.raises(ValueError, match="cplb_eval_device_shard"):
            sh.eval(torch.from_numpy(x).to(cuda_device))
        outs = []
        for s in range(len(devices)):
            dev, b, e = sh.GetShard(s, N)
            xd = torch.from_numpy(x[b:e]).to(torch.device("cuda", dev))
            outs.append(sh.eval_shard(s, xd, g=True, jac=True, cost=True, grad=True))
        for d in set(devices):
            torch.cuda.synchronize(d)
        want = o.eval_batch(x[::53], nthreads=4)
        got = {k: np.concatenate([r[k].cpu().numpy() for r in outs])[::53] for k in ("g", "jac", "cost", "grad")}
        assert_parity(got, want, o, f"eval_shard/{devices}", x[::53])
        with pytest.raises(ValueError, match="out of range"):
            sh.eval_shard(len(devices), torch.from_numpy(x[:4]).to(cuda_device))
