// tools/pyMOPS/bindings.cpp -- the `pyMOPS` Python module over the B200 engine.
//
// Same module name, class / enum / function names and call signatures as the reference's pybind11
// module (YosefQiu/MOPS tools/pyMOPS/bindings.cpp:19-476), so scripts written against it
// (tutorial/pyMOPSAPI.py, tutorial/streamLine.py ...) keep working for the hot path: grid and
// solution fed through the setters with numpy arrays, MOPS_Init/Begin/AddGridMesh/AddAttribute/End/
// ActiveAttribute, MOPS_GenerateSeedsPoints, MOPS_RunStreamLine / MOPS_RunPathLine (list of dicts),
// MOPS_RunRemapping (list of (H,W,4) arrays), the timing getters.  Differences: numpy arrays are
// taken with bulk copies instead of per-element loops, and two additional *Flat entry points return
// whole (N, M, 3) arrays for large particle counts (a Python list of 64 M dicts is not an option).
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <cstring>
#include <limits>
#include <memory>

#include "api/MOPS.h"

namespace py = pybind11;
using arr_d = py::array_t<double, py::array::c_style | py::array::forcecast>;

namespace {
std::vector<vec3> to_vec3(const arr_d& a, const char* what)
{
    if (a.ndim() != 2 || a.shape(1) != 3) throw std::runtime_error(std::string(what) + " must have shape (N, 3)");
    std::vector<vec3> v(static_cast<size_t>(a.shape(0)));
    if (!v.empty()) std::memcpy(v.data(), a.data(), v.size() * sizeof(vec3));
    return v;
}
py::tuple tup(const vec2& v) { return py::make_tuple(v.x(), v.y()); }
vec2 to_vec2(const py::tuple& t, const char* what)
{
    if (t.size() != 2) throw std::runtime_error(std::string(what) + " must be a tuple of size 2");
    return vec2{t[0].cast<double>(), t[1].cast<double>()};
}
py::array_t<double> lines_array(const std::vector<MOPS::TrajectoryLine>& lines, bool velocity)
{
    const py::ssize_t n = static_cast<py::ssize_t>(lines.size());
    const py::ssize_t per = n ? static_cast<py::ssize_t>(lines[0].points.size()) : 0;
    py::array_t<double> out({n, per, py::ssize_t(3)});
    double* d = out.mutable_data();
    for (py::ssize_t i = 0; i < n; ++i) {
        const auto& src = velocity ? lines[i].velocity : lines[i].points;
        std::memcpy(d + i * per * 3, src.data(), static_cast<size_t>(per) * sizeof(vec3));
    }
    return out;
}
} // namespace

PYBIND11_MODULE(pyMOPS, m)
{
    m.doc() = "pyMOPS (B200-native engine)";

    py::enum_<MOPS::AttributeFormat>(m, "AttributeFormat")
        .value("kDouble", MOPS::AttributeFormat::kDouble).value("kFloat", MOPS::AttributeFormat::kFloat)
        .value("kChar", MOPS::AttributeFormat::kChar).value("kVec3", MOPS::AttributeFormat::kVec3);
    py::enum_<MOPS::CalcPositionType>(m, "CalcPositionType")
        .value("kCenter", MOPS::CalcPositionType::kCenter).value("kVertx", MOPS::CalcPositionType::kVertx)
        .value("kPoint", MOPS::CalcPositionType::kPoint);
    py::enum_<MOPS::CalcAttributeType>(m, "CalcAttributeType")
        .value("kZonalMerimoal", MOPS::CalcAttributeType::kZonalMerimoal).value("kVelocity", MOPS::CalcAttributeType::kVelocity)
        .value("kZTop", MOPS::CalcAttributeType::kZTop).value("kTemperature", MOPS::CalcAttributeType::kTemperature)
        .value("kSalinity", MOPS::CalcAttributeType::kSalinity).value("kAll", MOPS::CalcAttributeType::kAll);
    py::enum_<MOPS::VisualizeType>(m, "VisualizeType")
        .value("kFixedLayer", MOPS::VisualizeType::kFixedLayer).value("kFixedDepth", MOPS::VisualizeType::kFixedDepth);
    py::enum_<MOPS::CalcDirection>(m, "CalcDirection")
        .value("kForward", MOPS::CalcDirection::kForward).value("kBackward", MOPS::CalcDirection::kBackward);
    py::enum_<MOPS::CalcMethodType>(m, "CalcMethodType")
        .value("kRK4", MOPS::CalcMethodType::kRK4).value("kEuler", MOPS::CalcMethodType::kEuler);
    py::enum_<MOPS::SaveType>(m, "SaveType").value("kVTI", MOPS::SaveType::kVTI).value("kNone", MOPS::SaveType::kNone);
    py::enum_<MOPS::GridAttributeType>(m, "GridAttributeType")
        .value("kCellSize", MOPS::GridAttributeType::kCellSize).value("kEdgeSize", MOPS::GridAttributeType::kEdgeSize)
        .value("kVertexSize", MOPS::GridAttributeType::kVertexSize).value("kMaxEdgesSize", MOPS::GridAttributeType::kMaxEdgesSize)
        .value("kVertLevels", MOPS::GridAttributeType::kVertLevels).value("kVertLevelsP1", MOPS::GridAttributeType::kVertLevelsP1)
        .value("kVertexCoord", MOPS::GridAttributeType::kVertexCoord).value("kCellCoord", MOPS::GridAttributeType::kCellCoord)
        .value("kEdgeCoord", MOPS::GridAttributeType::kEdgeCoord).value("kVertexLatLon", MOPS::GridAttributeType::kVertexLatLon)
        .value("kVerticesOnCell", MOPS::GridAttributeType::kVerticesOnCell).value("kVerticesOnEdge", MOPS::GridAttributeType::kVerticesOnEdge)
        .value("kCellsOnVertex", MOPS::GridAttributeType::kCellsOnVertex).value("kCellsOnCell", MOPS::GridAttributeType::kCellsOnCell)
        .value("kNumberVertexOnCell", MOPS::GridAttributeType::kNumberVertexOnCell).value("kCellsOnEdge", MOPS::GridAttributeType::kCellsOnEdge)
        .value("kEdgesOnCell", MOPS::GridAttributeType::kEdgesOnCell).value("kCellWeight", MOPS::GridAttributeType::kCellWeight);
    py::enum_<MOPS::AttributeType>(m, "AttributeType")
        .value("kZonalVelocity", MOPS::AttributeType::kZonalVelocity).value("kMeridionalVelocity", MOPS::AttributeType::kMeridionalVelocity)
        .value("kVelocity", MOPS::AttributeType::kVelocity).value("kNormalVelocity", MOPS::AttributeType::kNormalVelocity)
        .value("kZTop", MOPS::AttributeType::kZTop).value("kLayerThickness", MOPS::AttributeType::kLayerThickness)
        .value("kBottomDepth", MOPS::AttributeType::kBottomDepth);

    py::class_<MOPS::MPASOReader, std::shared_ptr<MOPS::MPASOReader>>(m, "MPASOReader")
        .def(py::init<>())
        .def_static("readGridData", &MOPS::MPASOReader::readGridData, py::arg("yaml_path"))
        .def_static("readSolData", &MOPS::MPASOReader::readSolData, py::arg("yaml_path"), py::arg("time_str"), py::arg("time_index") = 0);

    py::class_<MOPS::MPASOGrid, std::shared_ptr<MOPS::MPASOGrid>>(m, "MPASOGrid")
        .def(py::init<>())
        .def("init_from_reader", [](MOPS::MPASOGrid& self, const std::shared_ptr<MOPS::MPASOReader>& reader) { self.initGrid(reader.get()); })
        // read-back of the arrays (numpy copies), e.g. to check what a reader delivered
        .def("getArray", [](const MOPS::MPASOGrid& g, const std::string& name) -> py::object {
            auto v3 = [](const std::vector<vec3>& v) {
                py::array_t<double> a({static_cast<py::ssize_t>(v.size()), py::ssize_t(3)});
                if (!v.empty()) std::memcpy(a.mutable_data(), v.data(), v.size() * sizeof(vec3));
                return py::object(a);
            };
            auto sz = [](const std::vector<size_t>& v) {
                py::array_t<int64_t> a(static_cast<py::ssize_t>(v.size()));
                for (size_t i = 0; i < v.size(); ++i) a.mutable_data()[i] = static_cast<int64_t>(v[i]);
                return py::object(a);
            };
            if (name == "cellCoord") return v3(g.cellCoord_vec);
            if (name == "vertexCoord") return v3(g.vertexCoord_vec);
            if (name == "verticesOnCell") return sz(g.verticesOnCell_vec);
            if (name == "cellsOnCell") return sz(g.cellsOnCell_vec);
            if (name == "cellsOnVertex") return sz(g.cellsOnVertex_vec);
            if (name == "nEdgesOnCell") return sz(g.numberVertexOnCell_vec);
            if (name == "refBottomDepth") return py::object(py::array_t<double>(static_cast<py::ssize_t>(g.cellRefBottomDepth_vec.size()), g.cellRefBottomDepth_vec.data()));
            throw std::runtime_error("unknown grid array " + name);
        })
        .def_readonly("mCellsSize", &MOPS::MPASOGrid::mCellsSize)
        .def_readonly("mVertexSize", &MOPS::MPASOGrid::mVertexSize)
        .def_readonly("mMaxEdgesSize", &MOPS::MPASOGrid::mMaxEdgesSize)
        .def_readonly("mMeshName", &MOPS::MPASOGrid::mMeshName)
        .def("setGridAttribute", &MOPS::MPASOGrid::setGridAttribute)
        .def("setGridAttributesVec3", [](MOPS::MPASOGrid& self, MOPS::GridAttributeType type, arr_d arr) {
            self.setGridAttributesVec3(type, to_vec3(arr, "Input array"));
        })
        .def("setGridAttributesVec2", [](MOPS::MPASOGrid& self, MOPS::GridAttributeType type, arr_d arr) {
            if (arr.ndim() != 2 || arr.shape(1) != 2) throw std::runtime_error("Input array must have shape (N, 2)");
            std::vector<vec2> v(static_cast<size_t>(arr.shape(0)));
            if (!v.empty()) std::memcpy(v.data(), arr.data(), v.size() * sizeof(vec2));
            self.setGridAttributesVec2(type, v);
        })
        .def("setGridAttributesInt", [](MOPS::MPASOGrid& self, MOPS::GridAttributeType type,
                                        py::array_t<size_t, py::array::c_style | py::array::forcecast> arr) {
            if (arr.ndim() != 1) throw std::runtime_error("Input array must be 1D");
            self.setGridAttributesInt(type, std::vector<size_t>(arr.data(), arr.data() + arr.shape(0)));
        })
        .def("setRefBottomDepth", [](MOPS::MPASOGrid& self, arr_d arr) {
            self.cellRefBottomDepth_vec.assign(arr.data(), arr.data() + arr.size());
        })
        .def("setGridAttributesFloat", [](MOPS::MPASOGrid& self, MOPS::GridAttributeType type,
                                          py::array_t<float, py::array::c_style | py::array::forcecast> arr) {
            if (arr.ndim() != 1) throw std::runtime_error("Input array must be 1D");
            self.setGridAttributesFloat(type, std::vector<float>(arr.data(), arr.data() + arr.shape(0)));
        });

    py::class_<MOPS::MPASOSolution, std::shared_ptr<MOPS::MPASOSolution>>(m, "MPASOSolution")
        .def(py::init<>())
        .def("init_from_reader", [](MOPS::MPASOSolution& self, const std::shared_ptr<MOPS::MPASOReader>& reader) { self.initSolution(reader.get()); })
        .def("add_attribute", &MOPS::MPASOSolution::addAttribute)
        .def("getArray", [](const MOPS::MPASOSolution& s, const std::string& name) -> py::object {
            auto d = [](const std::vector<double>& v) { return py::object(py::array_t<double>(static_cast<py::ssize_t>(v.size()), v.data())); };
            if (name == "velocityZonal") return d(s.cellZonalVelocity_vec);
            if (name == "velocityMeridional") return d(s.cellMeridionalVelocity_vec);
            if (name == "layerThickness") return d(s.cellLayerThickness_vec);
            if (name == "bottomDepth") return d(s.cellBottomDepth_vec);
            if (name == "vertVelocityTop") return d(s.cellVertVelocity_vec);
            auto it = s.mDoubleAttributes.find(name);
            if (it != s.mDoubleAttributes.end()) return d(it->second);
            throw std::runtime_error("unknown solution array " + name);
        })
        .def_readonly("mVertLevels", &MOPS::MPASOSolution::mVertLevels)
        .def_readonly("mDataName", &MOPS::MPASOSolution::mDataName)
        .def("setTimestep", &MOPS::MPASOSolution::setTimestep)
        .def("setAttribute", &MOPS::MPASOSolution::setAttribute)
        .def("setAttributesDouble", [](MOPS::MPASOSolution& self, MOPS::AttributeType type, arr_d arr) {
            if (arr.ndim() != 1) throw std::runtime_error("Input must be (N,) numpy array");
            self.setAttributesDouble(type, std::vector<double>(arr.data(), arr.data() + arr.shape(0)));
        })
        // scalar tracers ("temperature", "salinity"): the reference fills mDoubleAttributes from the file
        // reader (add_attribute); without files they are handed over as arrays
        .def("setDoubleAttribute", [](MOPS::MPASOSolution& self, const std::string& name, arr_d arr) {
            self.mDoubleAttributes[name] = std::vector<double>(arr.data(), arr.data() + arr.size());
        })
        .def("setVertVelocityTop", [](MOPS::MPASOSolution& self, arr_d arr) {
            self.cellVertVelocity_vec.assign(arr.data(), arr.data() + arr.size());
        })
        .def_readwrite("mTimeStamp", &MOPS::MPASOSolution::mTimeStamp)
        .def("getID", &MOPS::MPASOSolution::getID)
        .def("getTimeStamp", &MOPS::MPASOSolution::getTimeStamp);

    py::class_<MOPS::VisualizationSettings>(m, "VisualizationSettings")
        .def(py::init<>())
        .def_property("imageSize", [](const MOPS::VisualizationSettings& s) { return tup(s.imageSize); },
                      [](MOPS::VisualizationSettings& s, py::tuple t) { s.imageSize = to_vec2(t, "imageSize"); })
        .def_property("LatRange", [](const MOPS::VisualizationSettings& s) { return tup(s.LatRange); },
                      [](MOPS::VisualizationSettings& s, py::tuple t) { s.LatRange = to_vec2(t, "LatRange"); })
        .def_property("LonRange", [](const MOPS::VisualizationSettings& s) { return tup(s.LonRange); },
                      [](MOPS::VisualizationSettings& s, py::tuple t) { s.LonRange = to_vec2(t, "LonRange"); })
        .def_property("DepthRange", [](const MOPS::VisualizationSettings& s) { return tup(s.DepthRange); },
                      [](MOPS::VisualizationSettings& s, py::tuple t) { s.DepthRange = to_vec2(t, "DepthRange"); })
        .def_readwrite("FixedLatitude", &MOPS::VisualizationSettings::FixedLatitude)
        .def_property("FixedLayer", [](const MOPS::VisualizationSettings& s) { return s.FixedLayer; },
                      [](MOPS::VisualizationSettings& s, double v) { s.FixedLayer = v; })
        .def_property("FixedDepth", [](const MOPS::VisualizationSettings& s) { return s.FixedDepth; },
                      [](MOPS::VisualizationSettings& s, double v) { s.FixedDepth = v; })
        .def_readwrite("TimeStep", &MOPS::VisualizationSettings::TimeStep)
        .def_readwrite("CalcType", &MOPS::VisualizationSettings::CalcType)
        .def_readwrite("VisType", &MOPS::VisualizationSettings::VisType)
        .def_readwrite("PositionType", &MOPS::VisualizationSettings::PositionType)
        .def_readwrite("SaveType", &MOPS::VisualizationSettings::saveType);

    py::class_<MOPS::SamplingSettings>(m, "SeedsSettings")
        .def(py::init<>())
        .def("setSeedsRange", [](MOPS::SamplingSettings& self, py::tuple t) {
            if (t.size() != 2) throw std::runtime_error("sampleRange must be a tuple of size 2");
            self.setSampleRange(vec2i{t[0].cast<int>(), t[1].cast<int>()});
        })
        .def("setGeoBox", [](MOPS::SamplingSettings& self, py::tuple lat, py::tuple lon) {
            self.setGeoBox(to_vec2(lat, "lat"), to_vec2(lon, "lon"));
        })
        .def("setDepth", &MOPS::SamplingSettings::setDepth)
        .def("getDepth", &MOPS::SamplingSettings::getDepth);

    py::class_<MOPS::TrajectorySettings>(m, "TrajectorySettings")
        .def(py::init<>())
        .def_readwrite("depth", &MOPS::TrajectorySettings::depth)
        .def_readwrite("particle_depths", &MOPS::TrajectorySettings::particle_depths)
        .def_readwrite("deltaT", &MOPS::TrajectorySettings::deltaT)
        .def_readwrite("simulationDuration", &MOPS::TrajectorySettings::simulationDuration)
        .def_readwrite("recordT", &MOPS::TrajectorySettings::recordT)
        .def_readwrite("directionType", &MOPS::TrajectorySettings::directionType)
        .def_readwrite("methodType", &MOPS::TrajectorySettings::methodType)
        .def_readwrite("fileName", &MOPS::TrajectorySettings::fileName)
        .def("hasPerParticleDepths", &MOPS::TrajectorySettings::hasPerParticleDepths);

    py::class_<CartesianCoord>(m, "CartesianCoord")
        .def(py::init<>())
        .def(py::init([](double x, double y, double z) { return CartesianCoord{x, y, z}; }))
        .def("x", [](const CartesianCoord& s) { return s.x(); })
        .def("y", [](const CartesianCoord& s) { return s.y(); })
        .def("z", [](const CartesianCoord& s) { return s.z(); });

    m.def("MOPS_Init", &MOPS::MOPS_Init, py::arg("device") = "gpu");
    m.def("MOPS_Begin", &MOPS::MOPS_Begin);
    m.def("MOPS_End", &MOPS::MOPS_End);
    m.def("MOPS_AddGridMesh", &MOPS::MOPS_AddGridMesh);
    m.def("MOPS_AddAttribute", &MOPS::MOPS_AddAttribute);
    m.def("MOPS_ActiveAttribute", &MOPS::MOPS_ActiveAttribute, py::arg("t1"), py::arg("t2") = py::none());

    m.def("MOPS_RunRemapping", [](MOPS::VisualizationSettings* config) {
        auto img_vec = MOPS::app.runRemapping(config);
        std::vector<py::array_t<double>> out;
        for (auto& img : img_vec)
            out.emplace_back(py::array_t<double>({img.getHeight(), img.getWidth(), 4}, img.mPixels.data())); // copies
        return out;
    });

    m.def("MOPS_RunReGrid", [](MOPS::VisualizationSettings* config) {
        auto img = MOPS::app.runReGrid(config);
        return py::array_t<double>({img.getHeight(), img.getWidth(), 4}, img.mPixels.data());
    }, "Run regridding at fixed latitude");
    m.def("MOPS_RunFixedLayer", [](MOPS::VisualizationSettings* config) {
        auto img = MOPS::app.runFixedLayer(config);
        return py::array_t<double>({img.getHeight(), img.getWidth(), 4}, img.mPixels.data());
    }, "Velocity of one layer (VisualizeFixedLayer)");
    m.def("MOPS_GenerateSeedsPoints", [](MOPS::SamplingSettings* setting) {
        std::vector<CartesianCoord> pts;
        MOPS::MOPS_GenerateSamplePoints(setting, pts);
        py::array_t<double> arr({static_cast<py::ssize_t>(pts.size()), py::ssize_t(3)});
        if (!pts.empty()) std::memcpy(arr.mutable_data(), pts.data(), pts.size() * sizeof(vec3));
        return arr;
    });

    m.def("MOPS_RunStreamLine", [](MOPS::TrajectorySettings* config, arr_d sample_points_np) {
        auto seeds = to_vec3(sample_points_np, "Input sample_points");
        auto lines = MOPS::MOPS_RunStreamLine(config, seeds);
        py::list out;
        for (const auto& ln : lines) {
            const py::ssize_t np_ = static_cast<py::ssize_t>(ln.points.size());
            py::array_t<double> pts({np_, py::ssize_t(3)}), vel({np_, py::ssize_t(3)});
            std::memcpy(pts.mutable_data(), ln.points.data(), static_cast<size_t>(np_) * sizeof(vec3));
            std::memcpy(vel.mutable_data(), ln.velocity.data(), static_cast<size_t>(np_) * sizeof(vec3));
            py::dict d;
            d["lineID"] = ln.lineID;
            d["points"] = std::move(pts);
            d["velocity"] = std::move(vel);
            out.append(std::move(d));
        }
        return out;
    }, "Run streamline simulation");

    m.def("MOPS_RunPathLine", [](MOPS::TrajectorySettings* config, arr_d sample_points_np) {
        auto seeds = to_vec3(sample_points_np, "Input sample_points");
        auto lines = MOPS::MOPS_RunPathLine(config, seeds);
        py::list out;
        for (const auto& ln : lines) {
            const py::ssize_t np_ = static_cast<py::ssize_t>(ln.points.size());
            py::array_t<double> pts({np_, py::ssize_t(3)}), vel({np_, py::ssize_t(3)}), tem(np_), sal(np_), last(3);
            std::memcpy(pts.mutable_data(), ln.points.data(), static_cast<size_t>(np_) * sizeof(vec3));
            std::memcpy(vel.mutable_data(), ln.velocity.data(), static_cast<size_t>(np_) * sizeof(vec3));
            std::memcpy(tem.mutable_data(), ln.temperature.data(), static_cast<size_t>(np_) * 8);
            std::memcpy(sal.mutable_data(), ln.salinity.data(), static_cast<size_t>(np_) * 8);
            std::memcpy(last.mutable_data(), &ln.lastPoint, sizeof(vec3));
            py::dict d;
            d["lineID"] = ln.lineID;
            d["points"] = std::move(pts);
            d["velocity"] = std::move(vel);
            d["temperature"] = std::move(tem);
            d["salinity"] = std::move(sal);
            d["lastPoint"] = std::move(last);
            d["depth"] = ln.depth;
            out.append(d);
        }
        return out;
    }, py::arg("config"), py::arg("sample_points_np"), "Run pathline simulation");

    // whole-array forms for large N: (points[N,M,3], velocity[N,M,3]); the pathline form also returns
    // the end points (the reference overwrites the caller's seeds, src/Core/MOPSApp.cpp:287-290)
    m.def("MOPS_RunStreamLineFlat", [](MOPS::TrajectorySettings* config, arr_d sample_points_np) {
        auto seeds = to_vec3(sample_points_np, "Input sample_points");
        auto lines = MOPS::MOPS_RunStreamLine(config, seeds);
        return py::make_tuple(lines_array(lines, false), lines_array(lines, true));
    });
    m.def("MOPS_RunPathLineFlat", [](MOPS::TrajectorySettings* config, arr_d sample_points_np) {
        auto seeds = to_vec3(sample_points_np, "Input sample_points");
        auto lines = MOPS::MOPS_RunPathLine(config, seeds);
        py::array_t<double> last({static_cast<py::ssize_t>(seeds.size()), py::ssize_t(3)});
        if (!seeds.empty()) std::memcpy(last.mutable_data(), seeds.data(), seeds.size() * sizeof(vec3));
        return py::make_tuple(lines_array(lines, false), lines_array(lines, true), last);
    });

    m.def("MOPS_ResetTiming", &MOPS::MOPS_ResetTiming, "Reset all timing data");
    m.def("MOPS_PrintTimingSummary", &MOPS::MOPS_PrintTimingSummary);
    m.def("MOPS_PrintTimingDetailed", &MOPS::MOPS_PrintTimingDetailed);
    m.def("MOPS_GetCategoryTime", &MOPS::MOPS_GetCategoryTime, py::arg("category"));
    m.def("MOPS_GetTotalTime", &MOPS::MOPS_GetTotalTime);
}
