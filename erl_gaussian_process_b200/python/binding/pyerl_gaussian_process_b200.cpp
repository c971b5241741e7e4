// pybind11 module over the drop-in host classes: the Python surface of the reference's pyerl_gaussian_process
// (python/binding/bind_vanilla_gp.cpp:79-101, bind_noisy_input_gp.cpp:134-182, bind_lidar_gp_2d.cpp:77-108,
// bind_range_sensor_gp_3d.cpp:80-126, bind_mapping.cpp) - same class names (VanillaGaussianProcessD / F, ...), nested Setting /
// TestResult classes, method names and argument names - with the numerics on the GPU through the C ABI.  Matrices cross the boundary
// as numpy arrays with the reference's shapes: x is (x_dim, n), y is (n, y_dim), directions are (3, T), column-major copies.
// Built by __graft_entry__.build() (g++ + pybind11, linked against liberl_gp_b200.so); tests/test_gpu_pybind.py drives it the way
// python/erl_gaussian_process users drive the reference.
#include "erl_gaussian_process_b200/lidar_gp_2d.hpp"
#include "erl_gaussian_process_b200/noisy_input_gp.hpp"
#include "erl_gaussian_process_b200/range_sensor_gp_3d.hpp"
#include "erl_gaussian_process_b200/vanilla_gp.hpp"

#include <pybind11/pybind11.h>
#include <pybind11/numpy.h>
#include <pybind11/stl.h>

#include <sstream>

namespace py = pybind11;
using namespace erl::gaussian_process;

namespace {

    template<typename T>
    using Array = py::array_t<T, py::array::f_style | py::array::forcecast>;

    template<typename T>
    Eigen::MatrixX<T>
    ToMatrix(const Array<T> &a) {  // (rows, cols) -> column-major copy; a 1-D array is one column
        const long rows = a.ndim() >= 1 ? a.shape(0) : 1, cols = a.ndim() >= 2 ? a.shape(1) : 1;
        Eigen::MatrixX<T> m(rows, cols);
        const T *src = a.data();
        for (long i = 0; i < rows * cols; ++i) { m.data()[i] = src[i]; }
        return m;
    }

    template<typename T>
    Eigen::VectorX<T>
    ToVector(const Array<T> &a) {
        const long n = a.size();
        Eigen::VectorX<T> v(n);
        for (long i = 0; i < n; ++i) { v[i] = a.data()[i]; }
        return v;
    }

    template<typename T, typename M>
    py::array_t<T>
    FromMatrix(const M &m) {
        py::array_t<T, py::array::f_style> out({static_cast<py::ssize_t>(m.rows()), static_cast<py::ssize_t>(m.cols())});
        for (long i = 0; i < m.rows() * m.cols(); ++i) { out.mutable_data()[i] = m.data()[i]; }
        return out;
    }

    template<typename T, typename V>
    py::array_t<T>
    FromVector(const V &v) {
        py::array_t<T> out(static_cast<py::ssize_t>(v.size()));
        for (long i = 0; i < v.size(); ++i) { out.mutable_data()[i] = v[i]; }
        return out;
    }

    template<typename V>
    py::array_t<bool>
    FromMask(const V &v) {
        py::array_t<bool> out(static_cast<py::ssize_t>(v.size()));
        for (long i = 0; i < v.size(); ++i) { out.mutable_data()[i] = v[i] != 0; }
        return out;
    }

    template<typename Dtype>
    void
    BindCovarianceSetting(const py::module &m, const char *name) {
        using S = typename erl::covariance::Covariance<Dtype>::Setting;
        py::class_<S, std::shared_ptr<S>>(m, name)
            .def(py::init<>())
            .def_readwrite("x_dim", &S::x_dim)
            .def_readwrite("scale", &S::scale)
            .def_readwrite("scale_mix", &S::scale_mix)
            .def_readwrite("weights", &S::weights);
    }

    template<typename Dtype>
    void
    BindVanilla(const py::module &m, const char *name) {
        using T = VanillaGaussianProcess<Dtype>;
        auto cls = py::class_<T, std::shared_ptr<T>>(m, name);
        py::class_<typename T::Setting, std::shared_ptr<typename T::Setting>>(cls, "Setting")
            .def(py::init<>())
            .def_readwrite("kernel_type", &T::Setting::kernel_type)
            .def_readwrite("kernel_setting_type", &T::Setting::kernel_setting_type)
            .def_readwrite("kernel", &T::Setting::kernel)
            .def_readwrite("max_num_samples", &T::Setting::max_num_samples);
        py::class_<typename T::TestResult, std::shared_ptr<typename T::TestResult>>(cls, "TestResult")
            .def_property_readonly("num_test", &T::TestResult::GetNumTest)
            .def(
                "get_mean",
                [](const typename T::TestResult &self, long y_index, bool parallel) {
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    self.GetMean(y_index, out, parallel);
                    return FromVector<Dtype>(out);
                },
                py::arg("y_index"),
                py::arg("parallel"))
            .def(
                "get_variance",
                [](const typename T::TestResult &self, bool parallel) {
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    self.GetVariance(out, parallel);
                    return FromVector<Dtype>(out);
                },
                py::arg("parallel"));
        cls.def(py::init([](std::shared_ptr<typename T::Setting> setting) { return std::make_shared<T>(std::move(setting)); }), py::arg("setting").none(false))
            .def_property_readonly("is_trained", &T::IsTrained)
            .def_property_readonly("setting", &T::GetSetting)
            .def("reset", &T::Reset)
            .def_property_readonly("k_train", [](const T &self) { return FromMatrix<Dtype>(self.GetKtrain()); })
            .def_property_readonly("alpha", [](const T &self) { return FromMatrix<Dtype>(self.GetAlpha()); })
            .def_property_readonly("cholesky_k_train", [](const T &self) { return FromMatrix<Dtype>(self.GetCholeskyDecomposition()); })
            .def(
                "train",
                [](T &self, const Array<Dtype> &mat_x_train, const Array<Dtype> &mat_y_train, const Array<Dtype> &vec_var_y) -> bool {  // bind_vanilla_gp.cpp:79-101
                    const long x_dim = mat_x_train.shape(0), n = mat_x_train.shape(1);
                    const long y_dim = mat_y_train.ndim() >= 2 ? mat_y_train.shape(1) : 1;
                    self.Reset(n, x_dim, y_dim);
                    auto &ts = self.GetTrainSet();
                    ts.x = ToMatrix<Dtype>(mat_x_train);
                    ts.y = ToMatrix<Dtype>(mat_y_train);
                    ts.var = ToVector<Dtype>(vec_var_y);
                    ts.x_dim = x_dim, ts.y_dim = y_dim, ts.num_samples = n;
                    return self.Train();
                },
                py::arg("mat_x_train"),
                py::arg("mat_y_train"),
                py::arg("vec_var_y"))
            .def(
                "test",
                [](const T &self, const Array<Dtype> &mat_x_test) { return self.Test(ToMatrix<Dtype>(mat_x_test)); },
                py::arg("mat_x_test"))
            .def("write", [](const T &self) {
                std::ostringstream s;
                if (!self.Write(s)) { throw std::runtime_error("Write failed"); }
                return py::bytes(s.str());
            })
            .def("read", [](T &self, const py::bytes &data) {
                std::istringstream s(static_cast<std::string>(data));
                return self.Read(s);
            })
            .def("__eq__", [](const T &a, const T &b) { return a == b; });
    }

    template<typename Dtype>
    void
    BindNoisy(const py::module &m, const char *name) {
        using T = NoisyInputGaussianProcess<Dtype>;
        auto cls = py::class_<T, std::shared_ptr<T>>(m, name);
        py::class_<typename T::Setting, std::shared_ptr<typename T::Setting>>(cls, "Setting")
            .def(py::init<>())
            .def_readwrite("kernel_type", &T::Setting::kernel_type)
            .def_readwrite("kernel_setting_type", &T::Setting::kernel_setting_type)
            .def_readwrite("kernel", &T::Setting::kernel)
            .def_readwrite("max_num_samples", &T::Setting::max_num_samples)
            .def_readwrite("no_gradient_observation", &T::Setting::no_gradient_observation);
        using R = typename T::TestResult;
        py::class_<R, std::shared_ptr<R>>(cls, "TestResult")
            .def_property_readonly("num_test", &R::GetNumTest)
            .def(
                "get_mean",
                [](const R &self, long y_index, bool parallel) {
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    self.GetMean(y_index, out, parallel);
                    return FromVector<Dtype>(out);
                },
                py::arg("y_index"),
                py::arg("parallel"))
            .def(
                "get_gradient",
                [](const R &self, long y_index, bool parallel) {  // -> (gradient (x_dim, num_test), valid)
                    Eigen::MatrixX<Dtype> out(self.GetDimX(), self.GetNumTest());
                    const auto valid = self.GetGradient(y_index, out, parallel);
                    return py::make_tuple(FromMatrix<Dtype>(out), FromMask(valid));
                },
                py::arg("y_index"),
                py::arg("parallel"))
            .def(
                "get_mean_variance",
                [](const R &self, bool parallel) {
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    self.GetMeanVariance(out, parallel);
                    return FromVector<Dtype>(out);
                },
                py::arg("parallel"))
            .def(
                "get_gradient_variance",
                [](const R &self, bool parallel) {
                    Eigen::MatrixX<Dtype> out(self.GetDimX(), self.GetNumTest());
                    self.GetGradientVariance(out, parallel);
                    return FromMatrix<Dtype>(out);
                },
                py::arg("parallel"))
            .def(
                "get_covariance",
                [](const R &self, bool parallel) {
                    Eigen::MatrixX<Dtype> out(self.GetDimX() * (self.GetDimX() + 1) / 2, self.GetNumTest());
                    self.GetCovariance(out, parallel);
                    return FromMatrix<Dtype>(out);
                },
                py::arg("parallel"));
        cls.def(py::init([](std::shared_ptr<typename T::Setting> setting) { return std::make_shared<T>(std::move(setting)); }), py::arg("setting").none(false))
            .def_property_readonly("setting", &T::GetSetting)
            .def_property_readonly("is_trained", &T::IsTrained)
            .def_property_readonly("using_reduced_rank_kernel", &T::UsingReducedRankKernel)
            .def("reset", &T::Reset, py::arg("max_num_samples"), py::arg("x_dim"), py::arg("y_dim"))
            .def_property_readonly("k_train", [](const T &self) { return FromMatrix<Dtype>(self.GetKtrain()); })
            .def_property_readonly("alpha", [](const T &self) { return FromMatrix<Dtype>(self.GetAlpha()); })
            .def_property_readonly("cholesky_k_train", [](const T &self) { return FromMatrix<Dtype>(self.GetCholeskyDecomposition()); })
            .def(
                "train",
                [](T &self, const Array<Dtype> &mat_x_train, const Array<Dtype> &mat_y_train, const Array<Dtype> &mat_grad_train, const Array<long> &vec_grad_flag,
                   const Array<Dtype> &vec_var_x, const Array<Dtype> &vec_var_y, const Array<Dtype> &vec_var_grad) {  // bind_noisy_input_gp.cpp:147-181
                    const long x_dim = mat_x_train.shape(0), n = mat_x_train.shape(1);
                    const long y_dim = mat_y_train.ndim() >= 2 ? mat_y_train.shape(1) : 1;
                    self.Reset(n, x_dim, y_dim);
                    auto &ts = self.GetTrainSet();
                    ts.x = ToMatrix<Dtype>(mat_x_train);
                    ts.y = ToMatrix<Dtype>(mat_y_train);
                    if (!self.GetSetting()->no_gradient_observation) {
                        ts.grad = ToMatrix<Dtype>(mat_grad_train);
                        ts.var_grad = ToVector<Dtype>(vec_var_grad);
                    }
                    ts.var_x = ToVector<Dtype>(vec_var_x);
                    ts.var_y = ToVector<Dtype>(vec_var_y);
                    ts.grad_flag = ToVector<long>(vec_grad_flag);
                    ts.x_dim = x_dim, ts.y_dim = y_dim, ts.num_samples = n;
                    long count = 0;
                    for (long i = 0; i < n; ++i) { count += ts.grad_flag[i] != 0; }
                    ts.num_samples_with_grad = self.GetSetting()->no_gradient_observation ? 0 : count;
                    return self.Train();
                },
                py::arg("mat_x_train"),
                py::arg("mat_y_train"),
                py::arg("mat_grad_train"),
                py::arg("vec_grad_flag"),
                py::arg("vec_var_x"),
                py::arg("vec_var_y"),
                py::arg("vec_var_grad"))
            .def(
                "test",
                [](const T &self, const Array<Dtype> &mat_x_test, bool predict_gradient) { return self.Test(ToMatrix<Dtype>(mat_x_test), predict_gradient); },
                py::arg("mat_x_test"),
                py::arg("predict_gradient"));
    }

    template<typename Dtype>
    void
    BindMapping(const py::module &m, const char *name) {
        using T = Mapping<Dtype>;
        auto cls = py::class_<T, std::shared_ptr<T>>(m, name);
        py::class_<typename T::Setting, std::shared_ptr<typename T::Setting>>(cls, "Setting")
            .def(py::init<>())
            .def_readwrite("type", &T::Setting::type)
            .def_readwrite("scale", &T::Setting::scale);
        cls.def(py::init([](std::shared_ptr<typename T::Setting> setting) { return T::Create(std::move(setting)); }), py::arg("setting"))
            .def("map", [](const T &self, Dtype x) { return self.map(x); })
            .def("inv", [](const T &self, Dtype y) { return self.inv(y); });
    }

    template<typename Dtype>
    void
    BindLidar(const py::module &m, const char *name) {
        using T = LidarGaussianProcess2D<Dtype>;
        using F = typename T::LidarFrameSetting;
        auto cls = py::class_<T, std::shared_ptr<T>>(m, name);
        py::class_<F, std::shared_ptr<F>>(cls, "SensorFrameSetting")
            .def(py::init<>())
            .def_readwrite("valid_range_min", &F::valid_range_min)
            .def_readwrite("valid_range_max", &F::valid_range_max)
            .def_readwrite("angle_min", &F::angle_min)
            .def_readwrite("angle_max", &F::angle_max)
            .def_readwrite("num_rays", &F::num_rays)
            .def_readwrite("discontinuity_detection", &F::discontinuity_detection);
        py::class_<typename T::Setting, std::shared_ptr<typename T::Setting>>(cls, "Setting")
            .def(py::init<>())
            .def_readwrite("partition_on_hit_rays", &T::Setting::partition_on_hit_rays)
            .def_readwrite("symmetric_partitions", &T::Setting::symmetric_partitions)
            .def_readwrite("group_size", &T::Setting::group_size)
            .def_readwrite("overlap_size", &T::Setting::overlap_size)
            .def_readwrite("margin", &T::Setting::margin)
            .def_readwrite("init_variance", &T::Setting::init_variance)
            .def_readwrite("sensor_range_var", &T::Setting::sensor_range_var)
            .def_readwrite("max_valid_range_var", &T::Setting::max_valid_range_var)
            .def_readwrite("occ_test_temperature", &T::Setting::occ_test_temperature)
            .def_readwrite("sensor_frame", &T::Setting::sensor_frame)
            .def_readwrite("gp", &T::Setting::gp)
            .def_readwrite("mapping", &T::Setting::mapping);
        using R = typename T::TestResult;
        py::class_<R, std::shared_ptr<R>>(cls, "TestResult")
            .def_property_readonly("num_test", &R::GetNumTest)
            .def(
                "get_mean",
                [](const R &self, bool parallel) {  // -> (success mask, mean); failed rays keep NaN (the reference leaves them unwritten)
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    for (long i = 0; i < out.size(); ++i) { out[i] = std::numeric_limits<Dtype>::quiet_NaN(); }
                    const auto ok = self.GetMean(out, parallel);
                    return py::make_tuple(FromMask(ok), FromVector<Dtype>(out));
                },
                py::arg("parallel"))
            .def(
                "get_variance",
                [](const R &self, bool parallel) {
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    for (long i = 0; i < out.size(); ++i) { out[i] = std::numeric_limits<Dtype>::quiet_NaN(); }
                    const auto ok = self.GetVariance(out, parallel);
                    return py::make_tuple(FromMask(ok), FromVector<Dtype>(out));
                },
                py::arg("parallel"));
        cls.def(py::init([](std::shared_ptr<typename T::Setting> setting) { return std::make_shared<T>(std::move(setting)); }), py::arg("setting").none(false))
            .def_property_readonly("is_trained", &T::IsTrained)
            .def_property_readonly("setting", &T::GetSetting)
            .def_property_readonly("angle_partitions", &T::GetAnglePartitions)
            .def_property_readonly("num_gps", [](const T &self) { return self.GetGps().size(); })
            .def("reset", &T::Reset)
            .def(
                "train",
                [](T &self, const Array<Dtype> &rotation, const Array<Dtype> &translation, const Array<Dtype> &ranges) {
                    return self.Train(ToMatrix<Dtype>(rotation), ToVector<Dtype>(translation), ToVector<Dtype>(ranges));
                },
                py::arg("rotation"),
                py::arg("translation"),
                py::arg("ranges"))
            .def(
                "test",
                [](const T &self, const Array<Dtype> &angles, bool angles_are_local, bool un_map) { return self.Test(ToVector<Dtype>(angles), angles_are_local, un_map); },
                py::arg("angles"),
                py::arg("angles_are_local"),
                py::arg("un_map"))
            .def(
                "compute_occ",
                [](const T &self, const Array<Dtype> &pos_local) {  // bind_lidar_gp_2d.cpp:96-108
                    Eigen::VectorX<Dtype> pos = ToVector<Dtype>(pos_local);
                    Dtype dist = 0, range_pred = 0, occ = 0;
                    const bool success = self.ComputeOcc(pos, dist, range_pred, occ);
                    py::dict out;
                    out["success"] = success;
                    out["dist_pos"] = dist;
                    out["range_pred"] = range_pred;
                    out["occ"] = occ;
                    return out;
                },
                py::arg("pos_local"));
    }

    template<typename Dtype>
    void
    BindRange3d(const py::module &m, const char *name) {
        using T = RangeSensorGaussianProcess3D<Dtype>;
        using F = typename T::RangeSensorFrame::Setting;
        auto cls = py::class_<T, std::shared_ptr<T>>(m, name);
        py::class_<F, std::shared_ptr<F>>(cls, "SensorFrameSetting")
            .def(py::init<>())
            .def_readwrite("valid_range_min", &F::valid_range_min)
            .def_readwrite("valid_range_max", &F::valid_range_max)
            .def_readwrite("azimuth_min", &F::azimuth_min)
            .def_readwrite("azimuth_max", &F::azimuth_max)
            .def_readwrite("elevation_min", &F::elevation_min)
            .def_readwrite("elevation_max", &F::elevation_max)
            .def_readwrite("num_azimuth_lines", &F::num_azimuth_lines)
            .def_readwrite("num_elevation_lines", &F::num_elevation_lines);
        py::class_<typename T::Setting, std::shared_ptr<typename T::Setting>>(cls, "Setting")
            .def(py::init<>())
            .def_readwrite("row_group_size", &T::Setting::row_group_size)
            .def_readwrite("row_overlap_size", &T::Setting::row_overlap_size)
            .def_readwrite("row_margin", &T::Setting::row_margin)
            .def_readwrite("col_group_size", &T::Setting::col_group_size)
            .def_readwrite("col_overlap_size", &T::Setting::col_overlap_size)
            .def_readwrite("col_margin", &T::Setting::col_margin)
            .def_readwrite("min_num_samples_per_group", &T::Setting::min_num_samples_per_group)
            .def_readwrite("init_variance", &T::Setting::init_variance)
            .def_readwrite("sensor_range_var", &T::Setting::sensor_range_var)
            .def_readwrite("max_valid_range_var", &T::Setting::max_valid_range_var)
            .def_readwrite("occ_test_temperature", &T::Setting::occ_test_temperature)
            .def_readwrite("sensor_frame_type", &T::Setting::sensor_frame_type)
            .def_readwrite("sensor_frame", &T::Setting::sensor_frame)
            .def_readwrite("gp", &T::Setting::gp)
            .def_readwrite("mapping", &T::Setting::mapping);
        using R = typename T::TestResult;
        py::class_<R, std::shared_ptr<R>>(cls, "TestResult")
            .def_property_readonly("num_test", &R::GetNumTest)
            .def(
                "get_mean",
                [](const R &self, bool parallel) {
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    for (long i = 0; i < out.size(); ++i) { out[i] = std::numeric_limits<Dtype>::quiet_NaN(); }
                    const auto ok = self.GetMean(out, parallel);
                    return py::make_tuple(FromMask(ok), FromVector<Dtype>(out));
                },
                py::arg("parallel"))
            .def(
                "get_variance",
                [](const R &self, bool parallel) {
                    Eigen::VectorX<Dtype> out(self.GetNumTest());
                    for (long i = 0; i < out.size(); ++i) { out[i] = std::numeric_limits<Dtype>::quiet_NaN(); }
                    const auto ok = self.GetVariance(out, parallel);
                    return py::make_tuple(FromMask(ok), FromVector<Dtype>(out));
                },
                py::arg("parallel"));
        cls.def(py::init([](std::shared_ptr<typename T::Setting> setting) { return std::make_shared<T>(std::move(setting)); }), py::arg("setting").none(false))
            .def_property_readonly("is_trained", &T::IsTrained)
            .def_property_readonly("setting", &T::GetSetting)
            .def_property_readonly("row_partitions", &T::GetRowPartitions)
            .def_property_readonly("col_partitions", &T::GetColPartitions)
            .def("reset", &T::Reset)
            .def(
                "train",
                [](T &self, const Array<Dtype> &rotation, const Array<Dtype> &translation, const Array<Dtype> &ranges) {
                    return self.Train(ToMatrix<Dtype>(rotation), ToVector<Dtype>(translation), ToMatrix<Dtype>(ranges));
                },
                py::arg("rotation"),
                py::arg("translation"),
                py::arg("ranges"))
            .def(
                "test",
                [](const T &self, const Array<Dtype> &directions, bool directions_are_local, bool un_map) { return self.Test(ToMatrix<Dtype>(directions), directions_are_local, un_map); },
                py::arg("directions"),
                py::arg("directions_are_local"),
                py::arg("un_map"))
            .def(
                "compute_occ",
                [](const T &self, const Array<Dtype> &pos_local) {  // bind_range_sensor_gp_3d.cpp:114-126
                    Eigen::VectorX<Dtype> pos = ToVector<Dtype>(pos_local);
                    Dtype dist = 0, range_pred = 0, occ = 0;
                    const bool success = self.ComputeOcc(pos, dist, range_pred, occ);
                    py::dict out;
                    out["success"] = success;
                    out["dist_pos"] = dist;
                    out["range_pred"] = range_pred;
                    out["occ"] = occ;
                    return out;
                },
                py::arg("pos_local"));
    }

}  // namespace

PYBIND11_MODULE(pyerl_gaussian_process_b200, m) {
    m.doc() = "Python 3 interface of erl_gaussian_process_b200 (the reference's pyerl_gaussian_process surface over the B200 C ABI)";
    py::enum_<MappingType>(m, "MappingType")
        .value("kIdentity", MappingType::kIdentity)
        .value("kInverse", MappingType::kInverse)
        .value("kInverseSqrt", MappingType::kInverseSqrt)
        .value("kExp", MappingType::kExp)
        .value("kLog", MappingType::kLog)
        .value("kTanh", MappingType::kTanh)
        .value("kSigmoid", MappingType::kSigmoid)
        .value("kUnknown", MappingType::kUnknown);
    BindCovarianceSetting<double>(m, "CovarianceSettingD");
    BindCovarianceSetting<float>(m, "CovarianceSettingF");
    BindVanilla<double>(m, "VanillaGaussianProcessD");
    BindVanilla<float>(m, "VanillaGaussianProcessF");
    BindMapping<double>(m, "MappingD");
    BindMapping<float>(m, "MappingF");
    BindLidar<double>(m, "LidarGaussianProcess2Dd");
    BindLidar<float>(m, "LidarGaussianProcess2Df");
    BindNoisy<double>(m, "NoisyInputGaussianProcessD");
    BindNoisy<float>(m, "NoisyInputGaussianProcessF");
    BindRange3d<double>(m, "RangeSensorGaussianProcess3Dd");
    BindRange3d<float>(m, "RangeSensorGaussianProcess3Df");
}
