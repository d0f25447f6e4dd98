"""The C++ compatibility classes (include/dlsc_compat.hpp: TrajPlanner / TrajOptimizer / CollisionConstraints
with the reference's signatures) driven the way MultiSyncSimulator drives the reference."""
import os
import re
import subprocess

import pytest

import _parity
from dlsc_gc_planner_b200 import capi

SRC = os.path.join(_parity.ROOT, "tests", "cpp", "compat_driver.cpp")
EXE = os.path.join(_parity.ROOT, "tests", "cpp", "compat_driver")


def build_driver():
    capi.build_library()
    libdir = os.path.dirname(capi.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(_parity.ROOT, "include"),
                           "-DDLSC_COMPAT_STANDALONE", SRC, "-o", EXE, "-L" + libdir, "-ldlsc_b200",
                           "-Wl,-rpath," + libdir])
    return EXE


def test_compat_header_compiles_and_links():
    """Signature check: the driver uses the reference's constructor / plan / setObstacles / solve signatures."""
    assert os.path.exists(build_driver())


def test_reference_signatures_are_mirrored():
    """Spot-check the declarations against the reference's own headers (quoted in SURVEY.md s8(b))."""
    hdr = open(os.path.join(_parity.ROOT, "include", "dlsc_compat.hpp")).read()
    for decl in (
        r"TrajPlanner\(const ros::NodeHandle& nh_, const Param& param_, const Mission& mission_, const Agent& agent_\)",
        r"TrajOptResult plan\(const Agent& agent_, const std::shared_ptr<octomap::OcTree>&[^,]*,\s*const std::shared_ptr<DynamicEDTOctomap>& distmap_ptr, ros::Time[^,]*,\s*bool is_disturbed\)",
        r"void setObstacles\(const Obstacles& obstacles_\)",
        r"int getPlannerSeq\(\) const", r"point3d getCurrentGoalPosition\(\) const",
        r"TrajOptimizer\(const Param& param_, const Mission& mission_, const Eigen::MatrixXd&\)",
        r"TrajOptResult solve\(const Agent& agent, const CollisionConstraints& constraints, const traj_t& initial_traj,",
        r"CollisionConstraints\(const Param& param_, const Mission& mission_, double radius, double max_vel\)",
        r"LSC getLSC\(int oi, int m, int i\) const", r"Box getSFC\(int m\) const",
    ):
        assert re.search(decl, hdr), decl


@pytest.mark.gpu
def test_compat_driver_serial_equals_batched():
    """plan() agent by agent (the unmodified reference loop) and the one-launch-per-step staged mode give
    bit-identical trajectories; agents stay collision free; TrajOptimizer::solve honours an explicit LSC."""
    exe = build_driver()
    outs = {}
    for mode in ("serial", "staged"):
        outs[mode] = subprocess.check_output([exe, mode, "25"], text=True).strip().split("\n")
    assert len(outs["serial"]) == 30
    for a, b in zip(outs["serial"][:25], outs["staged"][:25]):
        assert a == b, (a, b)
    st = re.search(r"stats seq (\d+) total (\S+) qp (\S+) samples (\d+)", outs["serial"][25])
    assert int(st.group(1)) == 25 and int(st.group(4)) == 25 and 0 < float(st.group(3)) <= float(st.group(2)) < 1.0
    assert outs["serial"][26] == "contexts distinct 1 same 1"
    dyn = re.search(r"dynamic obstacle planned cost (\S+) vs (\S+) end_y (\S+) vs (\S+) alerts (\d+) \(id (-?\d+)\) vs (\d+)", outs["serial"][27])
    assert float(dyn.group(1)) > float(dyn.group(2)) + 1e-3, outs["serial"][27]      # the obstacle costs something ...
    assert abs(float(dyn.group(3)) - float(dyn.group(4))) > 1e-3, outs["serial"][27]    # ... and bends the trajectory
    assert int(dyn.group(7)) == 0 and int(dyn.group(5)) in (0, 1) and (int(dyn.group(5)) == 0 or int(dyn.group(6)) == 0)
    assert outs["serial"][28] == "caller list kept 1"
    for line in outs["serial"][:25]:
        assert float(re.search(r"min_dist ([0-9.]+)", line).group(1)) >= 0.3 - 1e-4, line
    assert int(re.search(r"seq (\d+)", outs["serial"][24]).group(1)) == 25
    end = [float(x) for x in re.search(r"solve end (\S+) (\S+) (\S+)", outs["serial"][-1]).groups()]
    assert 1.9 - 1e-6 <= end[0] <= 1.9 + 1e-4 and abs(end[1]) < 1e-5 and abs(end[2] - 1.0) < 1e-5


REALTREE = os.path.join(_parity.ROOT, "tests", "cpp", "realtree")
SIG = os.path.join(_parity.ROOT, "tests", "cpp", "signature_check.cpp")


def test_member_signatures_are_exactly_the_references():
    """Every public member of the three classes has exactly the reference's type -- static_asserts on member
    pointers -- with the stand-in third-party types AND with headers that spell the real ros / octomap / dynamicEDT3D / Eigen
    APIs (tests/cpp/realtree), which also compiles the live-DynamicEDTOctomap export."""
    inc = "-I" + os.path.join(_parity.ROOT, "include")
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", inc, "-DDLSC_COMPAT_STANDALONE", SIG])
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", inc, "-I" + REALTREE, SIG])


@pytest.mark.gpu
def test_real_tree_map_binding(oracle):
    """Real-tree mode: the grid reaches the device through DynamicEDTOctomap::getDistanceAndClosestObstacle (one call per
    cell centre, SwarmBatch::set_distmap); a maze mission planned that way equals the same mission with the grid handed over
    directly through dlsc_set_edt."""
    import numpy as np
    from dlsc_gc_planner_b200 import missions
    capi.build_library()
    libdir = os.path.dirname(capi.LIB_PATH)
    exe = os.path.join(_parity.ROOT, "tests", "cpp", "realtree_driver")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + os.path.join(_parity.ROOT, "include"), "-I" + REALTREE, SIG, "-o", exe,
                           "-L" + libdir, "-ldlsc_b200", "-Wl,-rpath," + libdir])
    cfg, m = _parity.load_case("maze10")
    m = _parity.subset(m, 3)
    sw = _parity.make_oracle(cfg, m, 2)
    tmp = os.path.join(_parity.ROOT, "tests", "cpp")
    fd, fo = os.path.join(tmp, "_dist.bin"), os.path.join(tmp, "_obst.bin")
    sw.edt.dist.astype(np.float32).tofile(fd); sw.edt.obst.astype(np.int32).tofile(fo)
    start = m.start.copy(); start[:, 2] = cfg.z_2d
    goal = m.goal.copy(); goal[:, 2] = cfg.z_2d
    args = ([exe, fd, fo] + [str(x) for x in sw.edt.dims] + [str(x) for x in sw.edt.min_key] + ["3"] +
            [repr(float(x)) for x in m.world_min] + [repr(float(x)) for x in m.world_max])
    for a in range(3):
        args += [repr(float(x)) for x in start[a]] + [repr(float(x)) for x in goal[a]]
    out = subprocess.check_output(args, text=True).strip().split("\n")
    os.remove(fd); os.remove(fo)
    pl = capi.SwarmPlanner(cfg, m, max_nbr=2)
    pl.set_edt(sw.edt.dist, sw.edt.obst, sw.edt.dims, sw.edt.min_key, sw.edt.res)
    wts = (np.arange(1, cfg.M + 1)[:, None] * np.arange(1, cfg.n + 2)[None, :]).astype(np.float64)
    for step in range(3):
        pl.set_agents(waypoint=pl.start)
        pl.plan()
        t = pl.traj().astype(np.float64)
        checksum = float(sum((wts * (t[a, :, :, 0] + 2.0 * t[a, :, :, 1])).sum() for a in range(3)))
        got = float(re.search(r"checksum (\S+)", out[step]).group(1))
        assert abs(got - checksum) < 1e-6 * max(1.0, abs(checksum)), (step, got, checksum)
        pl.publish_records()
        pl.set_agents(pos=pl.traj()[:, 1, 0], vel=np.zeros((3, 3), np.float32), acc=np.zeros((3, 3), np.float32))
    pl.close()
