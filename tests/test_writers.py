"""CPU: dependency-free writers of the CLI: the PNG is a valid image (chunk CRCs, stored
deflate stream, Adler-32), TXT/VTK have the documented layout."""
import os
import struct
import subprocess
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r'''
#include "writers.hpp"
int main(int argc, char** argv) {
    MOPS::ImageBuffer<double> img(5, 3);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 5; ++j) img.setPixel(i, j, vec3(i, j, i * 5 + j));
    img.mPixels[(1 * 5 + 2) * 4 + 2] = std::nan("");
    MOPS::writers::SaveToPNG(img, std::string(argv[1]) + "/a.png", 2);
    std::vector<MOPS::TrajectoryLine> lines(2);
    for (int l = 0; l < 2; ++l) { lines[l].lineID = l; for (int k = 0; k < 3; ++k) { lines[l].points.push_back(vec3(l, k, 0.5)); lines[l].velocity.push_back(vec3(1, 2, 3)); } }
    MOPS::writers::SaveTrajectoryLinesAsTXT(lines, std::string(argv[1]) + "/a.txt");
    MOPS::writers::SaveTrajectoryLinesAsVTK(lines, std::string(argv[1]) + "/a.vtk");
    return 0;
}
'''


def test_writers(tmp_path):
    src = tmp_path / "w.cpp"
    src.write_text(SRC)
    exe = tmp_path / "w"
    subprocess.check_call(["g++", "-std=c++17", "-O1", f"-I{ROOT}/include", f"-I{ROOT}/mops_b200/host", "-o", str(exe), str(src)])
    subprocess.check_call([str(exe), str(tmp_path)])
    data = (tmp_path / "a.png").read_bytes()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, {}
    while pos < len(data):
        n = struct.unpack(">I", data[pos:pos + 4])[0]
        typ = data[pos + 4:pos + 8]
        body = data[pos + 8:pos + 8 + n]
        crc = struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0]
        assert zlib.crc32(typ + body) & 0xFFFFFFFF == crc
        chunks[typ] = body
        pos += 12 + n
    w, h, depth, ctype = struct.unpack(">IIBB", chunks[b"IHDR"][:10])
    assert (w, h, depth, ctype) == (5, 3, 8, 6)
    raw = zlib.decompress(chunks[b"IDAT"])          # validates the stored-deflate stream and its Adler-32
    px = np.frombuffer(raw, dtype=np.uint8).reshape(3, 1 + 5 * 4)[:, 1:].reshape(3, 5, 4)
    assert px[1, 2, 3] == 0 and (px[..., 3].sum() == 255 * 14)   # the NaN pixel is transparent
    assert tuple(px[0, 0, :3]) != tuple(px[2, 4, :3])
    txt = (tmp_path / "a.txt").read_text().splitlines()
    assert txt[0].startswith("Line_Index Point_Index") and txt[1] == "0 0 0 0 0.5 1 2 3" and len(txt) == 7
    vtk = (tmp_path / "a.vtk").read_text()
    assert "POINTS 6 double" in vtk and "LINES 2 8" in vtk and "VECTORS velocity double" in vtk
