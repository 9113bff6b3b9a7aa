"""Row-sharded local eigensolve (tnpy_eig_lowest_rows + the library's NCCL communicator) against the unsharded solver.
Runs scripts/sharded_solve.py under torchrun with one rank (always: the communicator path on a single GPU) and with
two ranks when the box has two GPUs."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_sharded(world, chi, extra=()):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29540 + world), os.path.join(ROOT, "scripts", "sharded_solve.py"), "--chi", str(chi), "--check",
           "--tol", "1e-10", *extra]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    return json.loads([line for line in res.stdout.splitlines() if line.startswith("{")][-1])


@pytest.mark.parametrize("world", [1, 2])
@pytest.mark.parametrize("chi,extra", [(64, ()), (64, ("--general",)), (256, ("--update",))])
def test_sharded_solve_matches_unsharded(world, chi, extra):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    out = run_sharded(world, chi, extra)
    assert out["converged"] and out["world"] == world
    assert out["theta_diff"] <= 1e-10 * abs(out["unsharded_theta"])
    assert out["max_abs_diff_psi_rows"] < 1e-7 and out["max_abs_diff_image_rows"] < 1e-6
    assert out["resid"] <= 1e-10 * 30
    if "--update" in extra:
        # the whole sharded local update (eigensolve + perturbation + gathered split + row-block environment update
        # summed over the ranks) against the product's own unsharded sweep step
        upd = out["local_update"]
        assert upd["theta_diff"] <= 1e-10 * abs(out["unsharded_theta"])
        assert upd["max_rel_diff_next_left_env"] < 1e-12
        assert upd["site_tensor_orthogonality_defect"] < 1e-12 and upd["two_site_state_rel_diff"] < 1e-12
