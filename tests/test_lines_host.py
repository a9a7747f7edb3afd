"""CPU: the C++ drop-in's line assembly (mops_b200/host/lines.hpp over mops_finalize_lines) -- the reference's
InitTrajectoryLines / FinalizeTrajectoryLines[WithAttrs] / RemoveNaNTrajectoriesAndReindex (src/Common/TrajectoryCommon.h:43-190)
-- against the oracle restatement, serial and on several host threads."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "lines.hpp"
#include <cstdio>
#include <fstream>
static std::vector<double> rd(const std::string& p) {
    std::ifstream f(p, std::ios::binary | std::ios::ate);
    std::vector<double> v(static_cast<size_t>(f.tellg()) / 8);
    f.seekg(0);
    f.read(reinterpret_cast<char*>(v.data()), static_cast<std::streamsize>(v.size() * 8));
    return v;
}
int main(int argc, char** argv) {
    const std::string dir = argv[1];
    const size_t n = std::stoul(argv[2]), each = std::stoul(argv[3]);
    const bool path = std::stoi(argv[4]) != 0;
    std::vector<double> seeds = rd(dir + "/seeds.bin"), rp = rd(dir + "/raw_pos.bin"), rv = rd(dir + "/raw_vel.bin");
    std::vector<float> depths(n);
    for (size_t i = 0; i < n; ++i) depths[i] = 10.0f + static_cast<float>(i);
    std::vector<MOPS::TrajectoryLine> lines = MOPS::detail::assemble_lines(
        n, each, reinterpret_cast<const CartesianCoord*>(seeds.data()), rp.data(), rv.data(), path, 86400.0, 120.0, depths.data());
    if (lines.size() != n) return 2;
    std::ofstream o(dir + "/out.bin", std::ios::binary);
    for (const auto& ln : lines) {
        if (ln.points.size() != each + 1 || ln.velocity.size() != each + 1 || ln.temperature.size() != each + 1 ||
            ln.salinity.size() != each + 1) return 3;
        const double head[4] = {static_cast<double>(ln.lineID), ln.duration, ln.timestamp, static_cast<double>(ln.depth)};
        o.write(reinterpret_cast<const char*>(head), 32);
        o.write(reinterpret_cast<const char*>(ln.points.data()), static_cast<std::streamsize>((each + 1) * 24));
        o.write(reinterpret_cast<const char*>(ln.velocity.data()), static_cast<std::streamsize>((each + 1) * 24));
        o.write(reinterpret_cast<const char*>(ln.temperature.data()), static_cast<std::streamsize>((each + 1) * 8));
        o.write(reinterpret_cast<const char*>(ln.salinity.data()), static_cast<std::streamsize>((each + 1) * 8));
        o.write(reinterpret_cast<const char*>(&ln.lastPoint), 24);
    }
    return 0;
}
'''


def test_cpp_line_assembly_matches_oracle(tmp_path):
    from oracle import port_oracle as P
    src = tmp_path / "lines.cpp"
    src.write_text(SRC)
    exe = tmp_path / "lines"
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT}/include", f"-I{ROOT}/mops_b200/host", "-o", str(exe), str(src),
                           f"-L{ROOT}/mops_b200", "-lmops_b200", "-pthread", f"-Wl,-rpath,{ROOT}/mops_b200",
                           "-Wl,-rpath,/usr/local/cuda/lib64"])
    rng = np.random.default_rng(17)
    n, each = 40000, 30  # 1.24 M points: the threaded path is taken
    per = each + 1
    seeds = rng.normal(size=(n, 3)); raw_pos = rng.normal(size=(n, each, 3)); raw_vel = rng.normal(size=(n, each, 3))
    raw_pos[::9, 12:, :] = np.nan
    raw_pos[3, 0, 1] = np.inf
    seeds[8, 0] = np.nan
    seeds.tofile(tmp_path / "seeds.bin"); raw_pos.tofile(tmp_path / "raw_pos.bin"); raw_vel.tofile(tmp_path / "raw_vel.bin")
    rec = 4 + per * 3 * 2 + per * 2 + 3
    for mode in (0, 1):
        ref = P.finalize_lines(seeds, raw_pos, raw_vel, pathline_mode=bool(mode))
        for threads in ("1", "5"):
            env = dict(os.environ, MOPS_HOST_THREADS=threads)
            subprocess.check_call([str(exe), str(tmp_path), str(n), str(each), str(mode)], env=env)
            out = np.fromfile(tmp_path / "out.bin").reshape(n, rec)
            assert np.array_equal(out[:, 0], np.arange(n))                       # lineID = input index
            assert (out[:, 1] == 86400.0).all() and (out[:, 2] == 120.0).all()
            assert np.array_equal(out[:, 3], 10.0 + np.arange(n))                 # depth the call started with
            o = 4
            pts = out[:, o:o + per * 3].reshape(n, per, 3); o += per * 3
            vel = out[:, o:o + per * 3].reshape(n, per, 3); o += per * 3
            temp = out[:, o:o + per]; o += per
            sal = out[:, o:o + per]; o += per
            last = out[:, o:o + 3]
            for got, key in ((pts, "points"), (vel, "velocity"), (temp, "temperature"), (sal, "salinity"), (last, "last")):
                assert np.array_equal(got, ref[key], equal_nan=True), (mode, threads, key)
