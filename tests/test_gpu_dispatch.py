"""Every dispatch target of the component-major launcher against the oracle.

`launch_cm_env` (csrc/cplb_kernels_cm.cuh) picks between two kernels by (contact count, batch size): one thread per
(instance, contact) and -- for shared-parameter 4- and 8-contact batches from a measured crossover on -- one thread per
instance.  Both are product code, so both are forced here (cplb_set_component_major_kernel) on the same inputs, for every
environment kind, both contact counts, every output subset, at the crossover, one below it and well above it, and must

  * agree bit for bit with the instance-major kernel (a third, independent code path) on EVERY instance, and
  * agree with the CPU oracle on a strided sample (bit-exact where no pow() is upstream, <= 1e-12 relative elsewhere).

Reference arithmetic these kernels restate: CentroidalStatics.cpp:44-57 (values), :121-135 (CoM block),
MinimizeCentroidalVariables.cpp:126-147 (cost), :163-191 (gradient)."""
import numpy as np
import pytest

import centroidalplanner_b200 as cpl
from centroidalplanner_b200 import _cabi

from helpers import assert_parity, make_pair

pytestmark = pytest.mark.gpu

CROSSOVER = {4: 90112, 8: 49152}  # launch_cm_env: whole_from
SUBSETS = [dict(g=True, jac=True), dict(g=True, jac=True, cost=True, grad=True), dict(g=True, jac=False), dict(g=False, jac=True),
           dict(g=False, jac=False, cost=True), dict(g=False, jac=False, grad=True)]
CASES = [("noenv4", 4), ("ground4", 4), ("superquadric4", 4), ("noenv8", 8), ("ground8", 8), ("superquadric8", 8)]


def _sizes(nc):
    return [CROSSOVER[nc] - 1, CROSSOVER[nc], 131072, 1 << 20]


@pytest.mark.parametrize("which", ["crossover-1", "crossover", "131072", "1048576"])
@pytest.mark.parametrize("case,nc", CASES)
def test_both_component_major_kernels_match_oracle_and_instance_major(case, nc, which, cuda_device):
    import torch

    N = dict(zip(["crossover-1", "crossover", "131072", "1048576"], _sizes(nc)))[which]
    prob, o, gen = make_pair(case)  # rich parameters: distinct refs / weights / thresholds per contact
    x = gen(N)
    xd = torch.from_numpy(x).to(cuda_device)
    xt = xd.t().contiguous()
    ref = prob.eval(xd, g=True, jac=True, cost=True, grad=True, layout=cpl.INSTANCE_MAJOR)  # every instance
    torch.cuda.synchronize()
    sub = np.unique(np.concatenate([np.arange(0, N, max(1, N // 1024)), [N - 1, N - 2, N - 33]]))
    want = o.eval_batch(x[sub], nthreads=4)
    subd = torch.from_numpy(sub).to(cuda_device)
    # the instance-major kernel itself against the oracle on the sample (it is the all-instance yardstick below)
    assert_parity({k: ref[k][subd].cpu().numpy() for k in ref}, want, o, f"{case}/N{N}/instance-major", x[sub])

    for kernel in ("split", "whole"):
        prob.SetComponentMajorKernel(kernel)
        for want_flags in SUBSETS:
            launches = prob.launch_count()
            out = prob.eval(xt, layout=cpl.COMPONENT_MAJOR, **want_flags)
            torch.cuda.synchronize()
            assert prob.launch_count() == launches + 1
            tag = f"{case}/N{N}/{kernel}/{'+'.join(k for k in want_flags if want_flags[k])}"
            got = {}
            for k in ("g", "jac", "cost", "grad"):
                if not want_flags.get(k, False):
                    assert out[k] is None, tag
                    continue
                a = out[k] if k == "cost" else out[k].t().contiguous()
                assert torch.equal(a.view(torch.int64), ref[k].view(torch.int64)), f"{tag}: {k} differs from the instance-major kernel"
                got[k] = a[subd].cpu().numpy()
                del a
            assert_parity(got, {k: (want[k] if k in got else None) for k in want}, o, tag, x[sub])
            del out
    prob.SetComponentMajorKernel("auto")


@pytest.mark.parametrize("case", ["ground4", "ground8"])
def test_forced_whole_with_per_instance_parameters_stays_on_the_per_contact_kernel(case, cuda_device):
    """Per-instance parameter arrays exist only in the per-contact kernel; a forced PER_INSTANCE choice must not drop them."""
    import torch

    prob, o, gen = make_pair(case)
    N = 2000
    x = gen(N)
    rng = np.random.default_rng(3)
    wrench = rng.uniform(-50, 50, (N, 6))
    prob.SetComponentMajorKernel("whole")
    out = prob.eval(torch.from_numpy(np.ascontiguousarray(x.T)).to(cuda_device), g=True, jac=False, layout=cpl.COMPONENT_MAJOR,
                    per_instance={"wrench": torch.from_numpy(np.ascontiguousarray(wrench.T)).to(cuda_device)})
    torch.cuda.synchronize()
    g = out["g"].t().cpu().numpy()
    for i in (0, 1, N - 1):
        o.set_wrench(wrench[i])
        assert np.array_equal(g[i], o.eval(x[i], want=("g",))["g"])


def test_two_threads_with_different_output_subsets_share_one_kernel_instantiation(cuda_device):
    """The instance-major launcher raises a kernel's opt-in shared-memory limit per (kernel, device) and never lowers it:
    a thread asking for the Jacobian only (64 KB of tiles per CTA) and a thread asking for the constraint values only
    (27 KB) use the SAME kernel instantiation; a limit lowered by the second would make the first one's next launch fail."""
    import threading

    import torch

    prob, o, gen = make_pair("ground4")
    N = 8192
    x = gen(N)
    xd = torch.from_numpy(x).to(cuda_device)
    ref = prob.eval(xd, g=True, jac=True, layout=cpl.INSTANCE_MAJOR)
    torch.cuda.synchronize()
    errors, results = [], {}

    def work(key):
        try:
            st = torch.cuda.Stream(cuda_device)
            with torch.cuda.stream(st):
                for _ in range(200):
                    out = prob.eval(xd, g=(key == "g"), jac=(key == "jac"), layout=cpl.INSTANCE_MAJOR, stream=st.cuda_stream)
                st.synchronize()
            results[key] = out[key]
        except Exception as e:  # noqa: BLE001
            errors.append((key, e))

    th = [threading.Thread(target=work, args=(k,)) for k in ("jac", "g", "jac", "g")]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errors, errors
    for k in ("g", "jac"):
        assert torch.equal(results[k].view(torch.int64), ref[k].view(torch.int64))
