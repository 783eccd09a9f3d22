"""CPLB_JAC_PACKED: instance-major evaluations whose Jacobian slice holds only the x-dependent slots.

The packed values must be the same bits as the same slots of a full evaluation; cplb_unpack_jacobian must rebuild the full
values[] rows (constants from cplb_get_jacobian_constants: CentroidalStatics.cpp:93-95, EnvironmentNormal.cpp:66-68,
Ground.cpp:33-34,49) that equal the oracle's."""
import ctypes as C

import numpy as np
import pytest

import centroidalplanner_b200 as cpl
from centroidalplanner_b200 import _cabi

from helpers import CASES, assert_parity, make_pair, same_bits

pytestmark = pytest.mark.gpu

SHAPES = ["ground4", "superquadric4", "noenv4", "ground8", "noenv8", "superquadric8", "ground1", "superquadric3", "ground5", "ground12",
          "superquadric32", "superquadric4_fracP"]


@pytest.mark.parametrize("case", SHAPES)
def test_packed_slices_equal_the_full_evaluation_and_unpack_to_the_oracle(case, cuda_device):
    import torch

    prob, o, gen = make_pair(case)
    pmap = prob.GetPackedJacobianMap()
    mask, cval = prob.GetJacobianConstants()
    assert np.array_equal(pmap, np.nonzero(~mask)[0])          # the x-dependent slots, in slot order
    for N in (1, 2, 31, 32, 33, 100, 4097, 70001):
        x = gen(N)
        xd = torch.from_numpy(x).to(cuda_device)
        full = prob.eval(xd, g=True, jac=True, cost=True, grad=True)
        for want in (dict(g=True, jac=True), dict(g=False, jac=True), dict(g=True, jac=True, cost=True, grad=True)):
            out = prob.eval(xd, jac_packed=True, **want)
            torch.cuda.synchronize()
            assert tuple(out["jac"].shape) == (N, len(pmap))
            assert torch.equal(out["jac"].view(torch.int64), full["jac"][:, torch.from_numpy(pmap).to(cuda_device).long()].view(torch.int64)), (case, N)
            for key in ("g", "cost", "grad"):
                if want.get(key, False):
                    assert torch.equal(out[key].view(torch.int64), full[key].view(torch.int64)), (case, N, key)
                else:
                    assert out[key] is None
        if N <= 4097:
            unpacked = prob.UnpackJacobian(out["jac"].cpu().numpy())
            assert same_bits(unpacked, full["jac"].cpu().numpy())
            want_o = o.eval_batch(x, nthreads=4)
            assert_parity({"g": out["g"].cpu().numpy(), "jac": unpacked, "cost": out["cost"].cpu().numpy(), "grad": out["grad"].cpu().numpy()},
                          want_o, o, f"packed/{case}/N{N}", x)


def test_packed_needs_instance_major(cuda_device):
    import torch

    prob, o, gen = make_pair("ground4")
    x = torch.from_numpy(np.ascontiguousarray(gen(64).T)).to(cuda_device)
    with pytest.raises(ValueError, match="INSTANCE_MAJOR"):
        prob.eval(x, layout=cpl.COMPONENT_MAJOR, jac_packed=True)
    with pytest.raises(ValueError, match="INSTANCE_MAJOR"):
        prob.eval(np.ascontiguousarray(gen(64).T), layout=cpl.COMPONENT_MAJOR, jac_packed=True)


@pytest.mark.parametrize("case", ["ground4", "noenv8", "superquadric3"])
def test_packed_with_per_instance_parameters(case, cuda_device):
    import torch

    prob, o, gen = make_pair(case)
    N, nc = 777, o.nc
    x = gen(N)
    rng = np.random.default_rng(11)
    pi = {"mass": rng.uniform(20, 150, N), "wrench": rng.uniform(-50, 50, (N, 6)), "mu": rng.uniform(0.2, 1.2, N),
          "force_threshold": rng.uniform(0, 30, (N, nc))}
    pid = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in pi.items()}
    xd = torch.from_numpy(x).to(cuda_device)
    full = prob.eval(xd, g=True, jac=True, per_instance=pid)
    out = prob.eval(xd, g=True, jac=True, per_instance=pid, jac_packed=True)
    torch.cuda.synchronize()
    pmap = torch.from_numpy(prob.GetPackedJacobianMap()).to(cuda_device).long()
    assert torch.equal(out["jac"].view(torch.int64), full["jac"][:, pmap].view(torch.int64))
    assert torch.equal(out["g"].view(torch.int64), full["g"].view(torch.int64))


@pytest.mark.parametrize("mode", ["pinned", "pageable", "queued", "sharded"])
@pytest.mark.parametrize("case", ["ground4", "superquadric4", "noenv8"])
def test_packed_host_path(case, mode, cuda_device):
    """Host buffers: the packed slices travel instead of the full rows (35 % fewer device -> host bytes for ground4)."""
    lib = _cabi.load()
    prob, o, gen = make_pair(case)
    if mode == "sharded":
        from test_gpu_sharded_abi import sharded_twin

        _, prob, o, gen = sharded_twin(case, [0, 0])
    N = 70003
    x = gen(N)
    full = make_pair(case)[0].eval(x, g=True, jac=True)
    pmap = prob.GetPackedJacobianMap()
    nv = len(pmap)
    keep = []

    def pinned(shape):
        ptr = C.c_void_p()
        assert lib.cplb_host_alloc(int(np.prod(shape)) * 8, C.byref(ptr)) == 0
        keep.append(ptr)
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(int(np.prod(shape)),)).reshape(shape)
        a[...] = np.nan
        return a

    if mode == "pageable":
        got = prob.eval(x, g=True, jac=True, jac_packed=True)
    else:
        hx, hg, hj = pinned((N, o.n)), pinned((N, o.m)), pinned((N, nv))
        hx[...] = x
        if mode == "queued":
            ticket, got = prob.eval_host_begin(hx, {"g": hg, "jac": hj}, g=True, jac=True, jac_packed=True)
            prob.eval_host_wait(ticket)
        else:
            got = prob.eval(hx, g=True, jac=True, out={"g": hg, "jac": hj}, jac_packed=True)
    assert got["jac"].shape == (N, nv)
    assert same_bits(got["jac"], full["jac"][:, pmap]) and same_bits(got["g"], full["g"])
    assert same_bits(prob.UnpackJacobian(got["jac"][:1000]), full["jac"][:1000])
    for ptr in keep:
        lib.cplb_host_free(ptr)


# ---- CPLB_JAC_COMPUTED: the slice also leaves out the slots that are copies +-x[col] --------------------------------------------

@pytest.mark.parametrize("case", SHAPES)
def test_computed_slices_equal_the_full_evaluation_and_expand_to_the_oracle(case, cuda_device):
    import torch

    prob, o, gen = make_pair(case)
    kind, src = prob.GetJacobianSlotSources()
    comp = np.nonzero(kind == _cabi.SLOT_COMPUTED)[0]
    compd = torch.from_numpy(comp).to(cuda_device).long()
    for N in (1, 2, 31, 32, 33, 100, 4097, 70001):
        x = gen(N)
        xd = torch.from_numpy(x).to(cuda_device)
        full = prob.eval(xd, g=True, jac=True, cost=True, grad=True)
        for want in (dict(g=True, jac=True), dict(g=False, jac=True), dict(g=True, jac=True, cost=True, grad=True)):
            out = prob.eval(xd, jac_packed="computed", **want)
            torch.cuda.synchronize()
            assert tuple(out["jac"].shape) == (N, len(comp))
            assert torch.equal(out["jac"].view(torch.int64), full["jac"][:, compd].view(torch.int64)), (case, N)
            for key in ("g", "cost", "grad"):
                if want.get(key, False):
                    assert torch.equal(out[key].view(torch.int64), full[key].view(torch.int64)), (case, N, key)
                else:
                    assert out[key] is None
        if N <= 4097:
            expanded = prob.ExpandJacobian(x, out["jac"].cpu().numpy())
            assert same_bits(expanded, full["jac"].cpu().numpy())
            want_o = o.eval_batch(x, nthreads=4)
            assert_parity({"g": out["g"].cpu().numpy(), "jac": expanded, "cost": out["cost"].cpu().numpy(), "grad": out["grad"].cpu().numpy()},
                          want_o, o, f"computed/{case}/N{N}", x)


def test_computed_flag_validation(cuda_device):
    import torch

    prob, o, gen = make_pair("ground4")
    x = torch.from_numpy(np.ascontiguousarray(gen(64).T)).to(cuda_device)
    with pytest.raises(ValueError, match="INSTANCE_MAJOR"):
        prob.eval(x, layout=cpl.COMPONENT_MAJOR, jac_packed="computed")
    # the two slice formats exclude each other (straight through the C ABI)
    lib = _cabi.load()
    xi = torch.from_numpy(gen(64)).to(cuda_device)
    g = torch.empty(64, o.m, dtype=torch.float64, device=cuda_device)
    j = torch.empty(64, o.nnz, dtype=torch.float64, device=cuda_device)
    args = _cabi.EvalArgs(64, _cabi.INSTANCE_MAJOR, _cabi.JAC_PACKED | _cabi.JAC_COMPUTED, 64, xi.data_ptr(), g.data_ptr(), j.data_ptr(), None, None, None)
    assert lib.cplb_eval_device(prob._h, C.byref(args), None) == _cabi.INVALID_ARGUMENT


@pytest.mark.parametrize("case", ["ground4", "noenv8", "superquadric3"])
def test_computed_with_per_instance_parameters(case, cuda_device):
    import torch

    prob, o, gen = make_pair(case)
    N, nc = 777, o.nc
    x = gen(N)
    rng = np.random.default_rng(12)
    pi = {"mass": rng.uniform(20, 150, N), "wrench": rng.uniform(-50, 50, (N, 6)), "mu": rng.uniform(0.2, 1.2, N),
          "force_threshold": rng.uniform(0, 30, (N, nc))}
    pid = {k: torch.from_numpy(np.ascontiguousarray(v)).to(cuda_device) for k, v in pi.items()}
    xd = torch.from_numpy(x).to(cuda_device)
    full = prob.eval(xd, g=True, jac=True, per_instance=pid)
    out = prob.eval(xd, g=True, jac=True, per_instance=pid, jac_packed="computed")
    torch.cuda.synchronize()
    kind, _ = prob.GetJacobianSlotSources()
    comp = torch.from_numpy(np.nonzero(kind == _cabi.SLOT_COMPUTED)[0]).to(cuda_device).long()
    assert torch.equal(out["jac"].view(torch.int64), full["jac"][:, comp].view(torch.int64))
    assert torch.equal(out["g"].view(torch.int64), full["g"].view(torch.int64))


@pytest.mark.parametrize("mode", ["pinned", "pageable", "queued", "sharded"])
@pytest.mark.parametrize("case", ["ground4", "superquadric4", "noenv8"])
def test_computed_host_path(case, mode, cuda_device):
    """Host buffers: only the computed slots travel back (g + 54 instead of g + 174 doubles per ground4 instance)."""
    lib = _cabi.load()
    prob, o, gen = make_pair(case)
    if mode == "sharded":
        from test_gpu_sharded_abi import sharded_twin

        _, prob, o, gen = sharded_twin(case, [0, 0])
    N = 70003
    x = gen(N)
    full = make_pair(case)[0].eval(x, g=True, jac=True)
    kind, _ = prob.GetJacobianSlotSources()
    comp = np.nonzero(kind == _cabi.SLOT_COMPUTED)[0]
    nv = len(comp)
    keep = []

    def pinned(shape):
        ptr = C.c_void_p()
        assert lib.cplb_host_alloc(int(np.prod(shape)) * 8, C.byref(ptr)) == 0
        keep.append(ptr)
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(int(np.prod(shape)),)).reshape(shape)
        a[...] = np.nan
        return a

    if mode == "pageable":
        got = prob.eval(x, g=True, jac=True, jac_packed="computed")
    else:
        hx, hg, hj = pinned((N, o.n)), pinned((N, o.m)), pinned((N, nv))
        hx[...] = x
        if mode == "queued":
            ticket, got = prob.eval_host_begin(hx, {"g": hg, "jac": hj}, g=True, jac=True, jac_packed="computed")
            prob.eval_host_wait(ticket)
        else:
            got = prob.eval(hx, g=True, jac=True, out={"g": hg, "jac": hj}, jac_packed="computed")
    assert got["jac"].shape == (N, nv)
    assert same_bits(got["jac"], full["jac"][:, comp]) and same_bits(got["g"], full["g"])
    assert same_bits(prob.ExpandJacobian(x[:1000], got["jac"][:1000]), full["jac"][:1000])
    for ptr in keep:
        lib.cplb_host_free(ptr)
