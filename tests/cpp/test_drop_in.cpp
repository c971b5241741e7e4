// C++ parity driver for the drop-in host classes (erl_gaussian_process_b200/cpp): written the way the
// reference's gtests drive the API (test/gtest/test_vanilla_gp.cpp:13-45, test_lidar_gp_2d.cpp:130-175) and
// checked against the CPU oracle.  TEST INFRASTRUCTURE: this is the only C++ translation unit that includes the
// oracle.  Needs a CUDA device; run by tests/test_gpu_cpp_dropin.py.
#include "erl_gaussian_process_b200/lidar_gp_2d.hpp"
#include "erl_gaussian_process_b200/noisy_input_gp.hpp"
#include "erl_gaussian_process_b200/range_sensor_gp_3d.hpp"
#include "erl_gaussian_process_b200/spgp_occupancy_map.hpp"
#include "erl_gaussian_process_b200/vanilla_gp.hpp"

#include "../../oracle/erl_gp_oracle.hpp"

#include <cstdio>
#include <sstream>
#include <random>

using namespace erl::gaussian_process;

static int g_failures = 0;

#define CHECK(cond, ...)                                      \
    do {                                                      \
        if (!(cond)) {                                        \
            ++g_failures;                                     \
            std::printf("FAIL %s:%d %s ", __FILE__, __LINE__, #cond); \
            std::printf(__VA_ARGS__);                         \
            std::printf("\n");                                \
        }                                                     \
    } while (0)

template<typename Dtype>
static void
TestVanillaSiso(const char *name, const double tol) {
    // test/gtest/test_vanilla_gp.cpp:13-45: RBF 1-D, l = 0.5, n = 100, y = sin(x), 200 test points
    constexpr long n = 100, n_test = 200;
    auto setting = std::make_shared<typename VanillaGaussianProcess<Dtype>::Setting>();
    setting->kernel->scale = 0.5;
    setting->kernel->x_dim = 1;
    setting->kernel_type = "erl::covariance::RadialBiasFunction1d";
    setting->max_num_samples = n;
    VanillaGaussianProcess<Dtype> gp(setting);
    CHECK(gp.Test(Eigen::MatrixX<Dtype>(1, 3)) == nullptr, "Test before Train must return nullptr");
    gp.Reset(n, 1, 1);
    auto &train_set = gp.GetTrainSet();
    erl_gp_oracle::VanillaGp<Dtype> ref;
    ref.kernel_type = erl_gp_oracle::kRadialBiasFunction;
    ref.scale = 0.5;
    ref.max_num_samples_setting = n;
    ref.Reset(n, 1, 1);
    for (long i = 0; i < n; ++i) {
        const Dtype x = Dtype(2 * M_PI * i / (n - 1));
        train_set.x(0, i) = x;
        train_set.y(i, 0) = std::sin(x);
        train_set.var[i] = Dtype(0.001);
        ref.x[i] = x, ref.y[i] = std::sin(x), ref.var[i] = Dtype(0.001);
    }
    train_set.num_samples = n;
    ref.num_samples = n;
    CHECK(gp.Train(), "Train");
    CHECK(!gp.Train(), "second Train without Reset must return false (src/vanilla_gp.cpp:511-514)");
    CHECK(ref.Train(), "oracle Train");
    Eigen::MatrixX<Dtype> x_test(1, n_test);
    for (long i = 0; i < n_test; ++i) { x_test(0, i) = Dtype(2 * M_PI * i / (n_test - 1)); }
    auto result = gp.Test(x_test);
    CHECK(result != nullptr, "Test");
    Eigen::VectorX<Dtype> mean(n_test), var(n_test);
    result->GetMean(0, mean, true);
    result->GetVariance(var, true);
    std::vector<Dtype> m_ref(n_test), v_ref(n_test);
    ref.Test(x_test.data(), 1, n_test, m_ref.data(), v_ref.data(), true);
    double em = 0, ev = 0, mae = 0;
    for (long i = 0; i < n_test; ++i) {
        em = std::max(em, std::abs(double(mean[i]) - double(m_ref[i])));
        ev = std::max(ev, std::abs(double(var[i]) - double(v_ref[i])));
        mae += std::abs(double(mean[i]) - std::sin(double(x_test(0, i)))) / n_test;
    }
    CHECK(em < tol && ev < tol, "mean err %.3e var err %.3e", em, ev);
    if (sizeof(Dtype) == 8) { CHECK(std::abs(mae - 0.00024246430481069056) < 1e-11, "KAT test_vanilla_gp.cpp:103 mae=%.17g", mae); }
    CHECK(mae < 3.0e-4, "reference threshold");
    // host materialisation of L / alpha
    const auto &mat_l = gp.GetCholeskyDecomposition();
    double el = 0;
    for (long c = 0; c < n; ++c) {
        for (long r = 0; r < n; ++r) { el = std::max(el, std::abs(double(mat_l(r, c)) - double(ref.mat_l[r + c * ref.ld]))); }
    }
    CHECK(el < tol * 10, "L err %.3e", el);
    CHECK(gp.GetLltInfo() == 0, "llt info");
    std::printf("%s %s: mean err %.2e, var err %.2e, mae %.6e\n", g_failures ? "----" : "PASS", name, em, ev, mae);
}

// NoisyInputGaussianProcess driven the way the reference's gtest drives it (test/gtest/test_noisy_input_gp.cpp:13-60, 366-430):
// Reset, fill the TrainSet (x, y, grad, var_x, var_y, var_grad, grad_flag, num_samples, num_samples_with_grad), Train, Test with
// gradient prediction; 2-D inputs, every second sample with a gradient observation; every TestResult output against the oracle.
template<typename Dtype>
static void
TestNoisyInput(const char *name, const double tol) {
    using Gp = NoisyInputGaussianProcess<Dtype>;
    constexpr long n = 160, n_test = 900, d = 2;
    auto setting = std::make_shared<typename Gp::Setting>();
    setting->kernel_type = "erl::covariance::RadialBiasFunction2d";
    setting->kernel->scale = Dtype(0.5);
    setting->kernel->x_dim = d;
    setting->max_num_samples = n;
    Gp gp(setting);
    std::mt19937 rng(17);
    std::uniform_real_distribution<double> uni(-1, 1);
    Eigen::MatrixX<Dtype> xt(d, n_test);
    for (long i = 0; i < n_test; ++i) { xt(0, i) = Dtype(uni(rng)), xt(1, i) = Dtype(uni(rng)); }
    CHECK(gp.Test(xt, true) == nullptr, "Test before Train must return nullptr");
    gp.Reset(n, d, 1);
    auto &ts = gp.GetTrainSet();
    erl_gp_oracle::NoisyInputGp<Dtype> ref;
    ref.kernel_type = erl_gp_oracle::kRadialBiasFunction, ref.scale = Dtype(0.5);
    ref.x_dim = d, ref.y_dim = 1, ref.num_samples = n;
    long ng = 0;
    for (long i = 0; i < n; ++i) {
        const double a = uni(rng), b = uni(rng);
        ts.x(0, i) = Dtype(a), ts.x(1, i) = Dtype(b);
        ts.y(i, 0) = Dtype(std::sin(2 * a) * std::cos(b));
        ts.grad(0, i) = Dtype(2 * std::cos(2 * a) * std::cos(b));
        ts.grad(1, i) = Dtype(-std::sin(2 * a) * std::sin(b));
        ts.var_x[i] = Dtype(0.01), ts.var_y[i] = Dtype(0.01), ts.var_grad[i] = Dtype(0.02);
        ts.grad_flag[i] = i % 2 == 0;
        ng += i % 2 == 0;
        ref.x.push_back(ts.x(0, i)), ref.x.push_back(ts.x(1, i));
        ref.y.push_back(ts.y(i, 0));
        ref.grad.push_back(ts.grad(0, i)), ref.grad.push_back(ts.grad(1, i));
        ref.var_x.push_back(ts.var_x[i]), ref.var_y.push_back(ts.var_y[i]), ref.var_grad.push_back(ts.var_grad[i]);
        ref.grad_flag.push_back(ts.grad_flag[i]);
    }
    ts.num_samples = n;
    ts.num_samples_with_grad = ng;
    CHECK(gp.Train(), "Train");
    CHECK(!gp.Train(), "second Train() without Reset must fail");
    CHECK(ref.Train(), "oracle Train");
    const long m = n + d * ng;
    CHECK(gp.GetKtrainSized().rows() == m && gp.GetCholeskyDecomposition().cols() == m && gp.GetAlphaSized().rows() == m, "K / L / alpha are m x m, m = n + d ng");
    double el = 0;
    for (long c = 0; c < m; ++c) {
        for (long r = c; r < m; ++r) { el = std::max(el, std::abs(double(gp.GetCholeskyDecomposition()(r, c)) - double(ref.mat_l[r + c * m]))); }
    }
    CHECK(el < (sizeof(Dtype) == 4 ? 5e-4 : 1e-10), "L err %.3e", el);
    auto result = gp.Test(xt, true);
    CHECK(result != nullptr, "Test");
    Eigen::VectorX<Dtype> mean(n_test), var(n_test);
    Eigen::MatrixX<Dtype> grad(d, n_test), gvar(d, n_test), cov(d * (d + 1) / 2, n_test);
    result->GetMean(0, mean, true);
    const auto valid = result->GetGradient(0, grad, true);
    result->GetMeanVariance(var, true);
    result->GetGradientVariance(gvar, true);
    result->GetCovariance(cov, true);
    std::vector<Dtype> m_r(n_test), g_r(d * n_test), v_r(n_test), gv_r(d * n_test), c_r(3 * n_test);
    ref.Test(xt.data(), n_test, true, m_r.data(), g_r.data(), v_r.data(), gv_r.data(), c_r.data());
    const double prior = 3.0 / 0.25;
    double em = 0, eg = 0, ev = 0, egv = 0, ec = 0, sm = 0, sg = 0;
    for (long i = 0; i < n_test; ++i) {
        CHECK(bool(valid[i]), "gradient %ld must be valid", i);
        sm = std::max(sm, std::abs(double(m_r[i])));
        em = std::max(em, std::abs(double(mean[i]) - double(m_r[i])));
        ev = std::max(ev, std::abs(double(var[i]) - double(v_r[i])));
        for (long j = 0; j < d; ++j) {
            sg = std::max(sg, std::abs(double(g_r[j + i * d])));
            eg = std::max(eg, std::abs(double(grad(j, i)) - double(g_r[j + i * d])));
            egv = std::max(egv, std::abs(double(gvar(j, i)) - double(gv_r[j + i * d])) / prior);
        }
        for (long j = 0; j < 3; ++j) { ec = std::max(ec, std::abs(double(cov(j, i)) - double(c_r[j + i * 3])) / prior); }
    }
    CHECK(em / sm < tol && eg / sg < tol && ev < tol && egv < tol && ec < tol, "mean %.2e grad %.2e var %.2e grad var %.2e cov %.2e", em / sm, eg / sg, ev, egv, ec);
    Dtype f1 = 0, g1[2] = {0, 0};
    result->GetMean(5, 0, f1);
    CHECK(result->GetGradient(5, 0, g1) && f1 == mean[5] && g1[0] == grad(0, 5) && g1[1] == grad(1, 5), "single-index accessors");
    std::printf("%s %s: mean %.2e grad %.2e var %.2e grad var %.2e cov %.2e\n", g_failures ? "----" : "PASS", name, em / sm, eg / sg, ev, egv, ec);
}

// Write / Read / operator== (src/vanilla_gp.cpp:561-790, src/noisy_input_gp.cpp:909-1160; the reference's gtests round-trip every GP
// through Serialization<>::Write / Read and assert gp == gp_read, e.g. test_noisy_input_gp.cpp:182-185): a trained GP goes through a
// stream into a fresh object whose device state is rebuilt; equal objects, identical predictions, a damaged stream is rejected.
template<typename Dtype>
static void
TestSerialization(const char *name) {
    using Gp = VanillaGaussianProcess<Dtype>;
    using Ngp = NoisyInputGaussianProcess<Dtype>;
    constexpr long n = 300, n_test = 500, d = 2;
    std::mt19937 rng(23);
    std::uniform_real_distribution<double> uni(-1, 1);
    Eigen::MatrixX<Dtype> xt(d, n_test);
    for (long i = 0; i < n_test; ++i) { xt(0, i) = Dtype(uni(rng)), xt(1, i) = Dtype(uni(rng)); }
    {
        auto setting = std::make_shared<typename Gp::Setting>();
        setting->kernel_type = "erl::covariance::Matern32<double, 2>";
        setting->kernel->scale = Dtype(0.4);
        setting->max_num_samples = n;
        Gp gp(setting);
        gp.Reset(n, d, 2);
        auto &ts = gp.GetTrainSet();
        for (long i = 0; i < n; ++i) {
            ts.x(0, i) = Dtype(uni(rng)), ts.x(1, i) = Dtype(uni(rng));
            ts.y(i, 0) = Dtype(std::sin(3 * double(ts.x(0, i)))), ts.y(i, 1) = Dtype(std::cos(2 * double(ts.x(1, i))));
            ts.var[i] = Dtype(0.01);
        }
        ts.num_samples = n;
        std::stringstream untrained;
        CHECK(gp.Write(untrained), "Write (untrained)");
        CHECK(gp.Train(), "Train");
        std::stringstream stream;
        CHECK(gp.Write(stream), "Write");
        auto setting2 = std::make_shared<typename Gp::Setting>();
        Gp gp2(setting2);
        CHECK(gp != gp2, "a fresh GP differs from a trained one");
        CHECK(gp2.Read(stream), "Read");
        CHECK(gp == gp2 && gp2.IsTrained() && !gp2.Train(), "gp == gp_read, trained, second Train() refused");
        Eigen::VectorX<Dtype> m1(n_test), m2(n_test), v1(n_test), v2(n_test);
        gp.Test(xt)->GetMean(1, m1, true), gp2.Test(xt)->GetMean(1, m2, true);
        gp.Test(xt)->GetVariance(v1, true), gp2.Test(xt)->GetVariance(v2, true);
        // the mean is bit-identical; the variance kernel combines its split reductions with atomics (few test tiles): equal up to rounding
        bool same = true;
        for (long i = 0; i < n_test; ++i) { same = same && m1[i] == m2[i] && std::abs(double(v1[i]) - double(v2[i])) < (sizeof(Dtype) == 4 ? 1e-5 : 1e-12); }
        CHECK(same, "predictions of the restored GP must match (mean bit for bit)");
        Gp gp3(std::make_shared<typename Gp::Setting>());
        CHECK(gp3.Read(untrained) && !gp3.IsTrained() && gp3.GetTrainSet() == gp.GetTrainSet() && gp3.Train() && gp3 == gp, "untrained round trip, then Train()");
        std::string bytes = stream.str();
        bytes[bytes.size() * 5 / 8] ^= 0x55;  // inside the stored L
        std::stringstream damaged(bytes);
        Gp gp4(std::make_shared<typename Gp::Setting>());
        CHECK(!gp4.Read(damaged) || gp4 != gp, "a damaged stream must not yield an equal GP");
    }
    {
        auto setting = std::make_shared<typename Ngp::Setting>();
        setting->kernel_type = "erl::covariance::RadialBiasFunction2d";
        setting->kernel->scale = Dtype(0.5);
        Ngp gp(setting);
        gp.Reset(n, d, 1);
        auto &ts = gp.GetTrainSet();
        long ng = 0;
        for (long i = 0; i < n; ++i) {
            const double a = uni(rng), b = uni(rng);
            ts.x(0, i) = Dtype(a), ts.x(1, i) = Dtype(b);
            ts.y(i, 0) = Dtype(std::sin(2 * a) * std::cos(b));
            ts.grad(0, i) = Dtype(2 * std::cos(2 * a) * std::cos(b)), ts.grad(1, i) = Dtype(-std::sin(2 * a) * std::sin(b));
            ts.var_x[i] = ts.var_y[i] = Dtype(0.01), ts.var_grad[i] = Dtype(0.02);
            ts.grad_flag[i] = i % 3 == 0;
            ng += i % 3 == 0;
        }
        ts.num_samples = n, ts.num_samples_with_grad = ng;
        CHECK(gp.Train(), "Train (noisy input)");
        std::stringstream stream;
        CHECK(gp.Write(stream), "Write (noisy input)");
        Ngp gp2(std::make_shared<typename Ngp::Setting>());
        CHECK(gp2.Read(stream) && gp == gp2, "NoisyInputGaussianProcess: gp == gp_read");  // test_noisy_input_gp.cpp:182-185
        Eigen::VectorX<Dtype> m1(n_test), m2(n_test);
        Eigen::MatrixX<Dtype> g1(d, n_test), g2(d, n_test);
        gp.Test(xt, true)->GetMean(0, m1, true), gp2.Test(xt, true)->GetMean(0, m2, true);
        (void) gp.Test(xt, true)->GetGradient(0, g1, true), (void) gp2.Test(xt, true)->GetGradient(0, g2, true);
        bool same = true;
        for (long i = 0; i < n_test; ++i) { same = same && m1[i] == m2[i] && g1(0, i) == g2(0, i) && g1(1, i) == g2(1, i); }
        CHECK(same, "predictions of the restored noisy-input GP must be bit-identical");
    }
    std::printf("%s %s\n", g_failures ? "----" : "PASS", name);
}

// Write / Read / operator== of the two sensor GPs (src/lidar_gp_2d.cpp:461-635, src/range_sensor_gp_3d.cpp:441-655)
template<typename Dtype>
static void
TestSensorSerialization(const char *name) {
    std::mt19937 rng(29);
    std::uniform_real_distribution<double> uni(0, 1);
    {
        using Lidar = LidarGaussianProcess2D<Dtype>;
        auto make_setting = [] {
            auto s = std::make_shared<typename Lidar::Setting>();
            s->group_size = 32, s->overlap_size = 8;
            s->sensor_frame->angle_min = Dtype(-2), s->sensor_frame->angle_max = Dtype(2), s->sensor_frame->num_rays = 360;
            s->sensor_frame->valid_range_min = Dtype(0.1), s->sensor_frame->valid_range_max = Dtype(30);
            s->gp->kernel_type = "erl::covariance::OrnsteinUhlenbeck1d";
            s->gp->kernel->scale = Dtype(0.05);
            return s;
        };
        Lidar gp(make_setting());
        Eigen::VectorX<Dtype> ranges(360);
        const auto &angles = gp.GetSensorFrame()->GetAnglesInFrame();
        for (long i = 0; i < 360; ++i) { ranges[i] = uni(rng) < 0.05 ? Dtype(1000) : Dtype(4 + std::sin(3 * double(angles[i]))); }
        Eigen::MatrixX<Dtype> rot(2, 2);
        rot.setZero();
        rot(0, 0) = rot(1, 1) = 1;
        Eigen::VectorX<Dtype> trans(2);
        trans.setZero();
        CHECK(gp.Train(rot, trans, ranges), "Train");
        std::stringstream stream;
        CHECK(gp.Write(stream), "LidarGaussianProcess2D::Write");
        Lidar gp2(make_setting());
        CHECK(gp != gp2, "untrained != trained");
        CHECK(gp2.Read(stream), "LidarGaussianProcess2D::Read");
        CHECK(gp == gp2 && gp2.IsTrained(), "LidarGaussianProcess2D: gp == gp_read");
        Eigen::VectorX<Dtype> q(500), m1(500), m2(500);
        for (long i = 0; i < 500; ++i) { q[i] = Dtype(-1.9 + 3.8 * uni(rng)), m1[i] = m2[i] = 0; }
        const auto ok1 = gp.Test(q, true, true)->GetMean(m1, true);
        const auto ok2 = gp2.Test(q, true, true)->GetMean(m2, true);
        bool same = true;
        for (long i = 0; i < 500; ++i) { same = same && bool(ok1[i]) == bool(ok2[i]) && m1[i] == m2[i]; }
        CHECK(same, "restored lidar GP predicts bit-identically");
        auto other = make_setting();
        other->overlap_size = 6;
        Lidar gp3(other);
        std::stringstream again(stream.str());
        CHECK(!gp3.Read(again), "a stream written under another Setting is rejected");
    }
    {
        using Rg = RangeSensorGaussianProcess3D<Dtype>;
        constexpr long rows = 40, cols = 48;
        auto make_setting = [] {
            auto s = std::make_shared<typename Rg::Setting>();
            s->row_group_size = 12, s->row_overlap_size = 2, s->col_group_size = 10, s->col_overlap_size = 4;
            s->sensor_frame->azimuth_min = Dtype(-0.4), s->sensor_frame->azimuth_max = Dtype(0.4), s->sensor_frame->num_azimuth_lines = rows;
            s->sensor_frame->elevation_min = Dtype(-0.3), s->sensor_frame->elevation_max = Dtype(0.3), s->sensor_frame->num_elevation_lines = cols;
            s->sensor_frame->valid_range_min = Dtype(0.1), s->sensor_frame->valid_range_max = Dtype(30);
            s->gp->kernel_type = "erl::covariance::Matern32<float, 2>";
            s->gp->kernel->scale = Dtype(0.05);
            return s;
        };
        Rg gp(make_setting());
        Eigen::MatrixX<Dtype> ranges(rows, cols), rot(3, 3);
        for (long c = 0; c < cols; ++c) {
            for (long r = 0; r < rows; ++r) { ranges(r, c) = uni(rng) < 0.05 ? Dtype(1000) : Dtype(4 + 0.5 * std::sin(r / 5.0) * std::cos(c / 7.0)); }
        }
        rot.setZero();
        rot(0, 0) = rot(1, 1) = rot(2, 2) = 1;
        Eigen::VectorX<Dtype> trans(3);
        trans.setZero();
        CHECK(gp.Train(rot, trans, ranges), "Train");
        std::stringstream stream;
        CHECK(gp.Write(stream), "RangeSensorGaussianProcess3D::Write");
        Rg gp2(make_setting());
        CHECK(gp2.Read(stream), "RangeSensorGaussianProcess3D::Read");
        CHECK(gp == gp2 && gp2.IsTrained(), "RangeSensorGaussianProcess3D: gp == gp_read");
    }
    std::printf("%s %s\n", g_failures ? "----" : "PASS", name);
}

// Setting::partition_on_hit_rays (src/lidar_gp_2d.cpp:302-348, 364): the table is empty after construction and follows the hit rays
// of every Train(); GetAnglePartitions() / GetGps() are refreshed.  Two frames through one object, against the oracle.
template<typename Dtype>
static void
TestLidarHitRays(const char *name, const double tol) {
    using Lidar = LidarGaussianProcess2D<Dtype>;
    constexpr long n = 540, n_test = 5000;
    auto setting = std::make_shared<typename Lidar::Setting>();
    setting->partition_on_hit_rays = true;
    setting->group_size = 40;
    setting->overlap_size = 10;
    setting->sensor_frame->angle_min = Dtype(-2.0);
    setting->sensor_frame->angle_max = Dtype(2.0);
    setting->sensor_frame->num_rays = n;
    setting->sensor_frame->valid_range_min = Dtype(0.1);
    setting->sensor_frame->valid_range_max = Dtype(30);
    setting->gp->kernel_type = "erl::covariance::OrnsteinUhlenbeck1d";
    setting->gp->kernel->scale = Dtype(0.05);
    Lidar gp(setting);
    CHECK(gp.GetAnglePartitions().empty() && gp.GetGps().empty(), "hit-ray partitions must be empty before Train");
    const auto &angles = gp.GetSensorFrame()->GetAnglesInFrame();
    erl_gp_oracle::LidarGp2D<Dtype> ref;
    ref.partition_on_hit_rays = true;
    ref.group_size = 40, ref.overlap_size = 10;
    ref.kernel_type = erl_gp_oracle::kOrnsteinUhlenbeck, ref.kernel_scale = Dtype(0.05);
    ref.sensor_range_var = setting->sensor_range_var;
    ref.Init(angles.data(), n);
    std::mt19937 rng(5);
    std::uniform_real_distribution<double> uni(0, 1);
    Eigen::MatrixX<Dtype> rot(2, 2);
    rot.setZero();
    rot(0, 0) = rot(1, 1) = 1;
    Eigen::VectorX<Dtype> trans(2);
    trans.setZero();
    double em = 0, ev = 0, scale = 0;
    for (const double miss: {0.05, 0.35}) {
        Eigen::VectorX<Dtype> ranges(n);
        for (long i = 0; i < n; ++i) {
            ranges[i] = Dtype(5 + 2 * std::sin(3 * double(angles[i])));
            if (uni(rng) < miss || i >= n - 3) { ranges[i] = Dtype(1000); }
        }
        CHECK(gp.Train(rot, trans, ranges), "Train (hit rays)");
        const auto frame = gp.GetSensorFrame();
        ref.Train(rot.data(), frame->GetRanges().data(), erl::gaussian_process::b200::MaskData(frame->GetHitMask()), erl::gaussian_process::b200::MaskData(frame->GetContinuityMask()), true);
        const auto &parts = gp.GetAnglePartitions();
        CHECK(parts.size() == ref.partitions.size() && gp.GetGps().size() == parts.size(), "hit-ray partitions: %zu vs %zu", parts.size(), ref.partitions.size());
        for (std::size_t i = 0; i < parts.size() && i < ref.partitions.size(); ++i) {
            CHECK(std::get<0>(parts[i]) == ref.partitions[i].index_left && std::get<1>(parts[i]) == ref.partitions[i].index_right, "partition %zu indices", i);
            CHECK(std::get<2>(parts[i]) == ref.partitions[i].coord_left && std::get<3>(parts[i]) == ref.partitions[i].coord_right, "partition %zu coords", i);
            CHECK(gp.GetGps()[i]->IsTrained() == ref.gps[i].trained, "partition GP %zu trained flag", i);
        }
        Eigen::VectorX<Dtype> q(n_test), mean(n_test), var(n_test);
        for (long i = 0; i < n_test; ++i) {
            q[i] = Dtype(-2.05 + uni(rng) * 4.1);
            mean[i] = var[i] = Dtype(-777);
        }
        auto result = gp.Test(q, true, true);
        CHECK(result != nullptr, "Test (hit rays)");
        const auto ok_mean = result->GetMean(mean, true);
        (void) result->GetVariance(var, true);
        std::vector<Dtype> m_ref(n_test, Dtype(-777)), v_ref(n_test, Dtype(-777));
        std::vector<uint8_t> ok_ref(n_test);
        ref.Test(q.data(), n_test, true, true, m_ref.data(), v_ref.data(), ok_ref.data());
        for (long i = 0; i < n_test; ++i) {
            CHECK(bool(ok_mean[i]) == bool(ok_ref[i]), "valid mask differs at %ld (hit rays)", i);
            if (!ok_ref[i]) { continue; }
            scale = std::max(scale, std::abs(double(m_ref[i])));
            em = std::max(em, std::abs(double(mean[i]) - double(m_ref[i])));
            ev = std::max(ev, std::abs(double(var[i]) - double(v_ref[i])));
        }
    }
    CHECK(em / scale < tol && ev < tol, "mean err %.3e var err %.3e", em / scale, ev);
    std::printf("%s %s: mean err %.2e, var err %.2e\n", g_failures ? "----" : "PASS", name, em / scale, ev);
}

template<typename Dtype>
static void
TestLidar(const char *name, const double tol) {
    // synthetic 1080-beam scan, group 64 / overlap 18 (BASELINE config 2), OU l = 0.05
    using Lidar = LidarGaussianProcess2D<Dtype>;
    constexpr long n = 1080, n_test = 20000;
    auto setting = std::make_shared<typename Lidar::Setting>();
    setting->group_size = 64;
    setting->overlap_size = 18;
    setting->margin = 1;
    setting->sensor_frame->angle_min = Dtype(-3 * M_PI / 4);
    setting->sensor_frame->angle_max = Dtype(3 * M_PI / 4);
    setting->sensor_frame->num_rays = n;
    setting->sensor_frame->valid_range_min = Dtype(0.1);
    setting->sensor_frame->valid_range_max = Dtype(30);
    setting->gp->kernel_type = "erl::covariance::OrnsteinUhlenbeck1d";
    setting->gp->kernel->scale = Dtype(0.05);
    Lidar gp(setting);
    CHECK(gp.GetAnglePartitions().size() == 24, "24 partitions expected, got %zu", gp.GetAnglePartitions().size());
    Eigen::VectorX<Dtype> dummy(1);
    dummy[0] = 0;
    CHECK(gp.Test(dummy, true, true) == nullptr, "Test before Train must return nullptr");

    std::mt19937 rng(3);
    std::uniform_real_distribution<double> uni(0, 1);
    Eigen::VectorX<Dtype> ranges(n);
    const auto &angles = gp.GetSensorFrame()->GetAnglesInFrame();
    for (long i = 0; i < n; ++i) {
        ranges[i] = Dtype(5 + 2 * std::sin(3 * double(angles[i])));
        if (uni(rng) < 0.02) { ranges[i] = Dtype(1000); }
    }
    Eigen::MatrixX<Dtype> rot(2, 2);
    rot.setZero();
    rot(0, 0) = rot(1, 1) = 1;
    Eigen::VectorX<Dtype> trans(2);
    trans.setZero();
    CHECK(gp.Train(rot, trans, ranges), "Train");
    const auto frame = gp.GetSensorFrame();

    erl_gp_oracle::LidarGp2D<Dtype> ref;
    ref.group_size = 64, ref.overlap_size = 18, ref.margin = 1;
    ref.kernel_type = erl_gp_oracle::kOrnsteinUhlenbeck, ref.kernel_scale = Dtype(0.05);
    ref.sensor_range_var = setting->sensor_range_var, ref.max_valid_range_var = setting->max_valid_range_var, ref.occ_test_temperature = setting->occ_test_temperature;
    ref.Init(angles.data(), n);
    ref.Train(rot.data(), frame->GetRanges().data(), erl::gaussian_process::b200::MaskData(frame->GetHitMask()), erl::gaussian_process::b200::MaskData(frame->GetContinuityMask()), true);

    Eigen::VectorX<Dtype> q(n_test), mean(n_test), var(n_test);
    for (long i = 0; i < n_test; ++i) {
        q[i] = Dtype(-3 * M_PI / 4 - 0.05 + uni(rng) * (1.5 * M_PI + 0.1));
        mean[i] = var[i] = Dtype(-777);
    }
    auto result = gp.Test(q, true, true);
    CHECK(result != nullptr, "Test");
    const auto ok_mean = result->GetMean(mean, true);
    const auto ok_var = result->GetVariance(var, true);
    std::vector<Dtype> m_ref(n_test, Dtype(-777)), v_ref(n_test, Dtype(-777));
    std::vector<uint8_t> ok_ref(n_test);
    ref.Test(q.data(), n_test, true, true, m_ref.data(), v_ref.data(), ok_ref.data());
    double em = 0, ev = 0, scale = 0;
    long invalid = 0;
    for (long i = 0; i < n_test; ++i) {
        CHECK(bool(ok_mean[i]) == bool(ok_ref[i]) && bool(ok_var[i]) == bool(ok_ref[i]), "valid mask differs at %ld", i);
        if (!ok_ref[i]) {
            ++invalid;
            CHECK(mean[i] == Dtype(-777) && var[i] == Dtype(-777), "invalid ray %ld must stay unwritten", i);
            continue;
        }
        scale = std::max(scale, std::abs(double(m_ref[i])));
        em = std::max(em, std::abs(double(mean[i]) - double(m_ref[i])));
        ev = std::max(ev, std::abs(double(var[i]) - double(v_ref[i])));
    }
    CHECK(invalid > 0, "expected some rays outside the field of view");
    CHECK(em / scale < tol && ev < tol, "mean err %.3e var err %.3e", em / scale, ev);
    // single-position ComputeOcc against the oracle
    long occ_checked = 0;
    for (int i = 0; i < 50; ++i) {
        const double a = -2.0 + 4.0 * uni(rng), d = 1.0 + 6.0 * uni(rng);
        Eigen::VectorX<Dtype> pos(2);
        pos[0] = Dtype(d * std::cos(a)), pos[1] = Dtype(d * std::sin(a));
        Dtype dist = 0, range_pred = 0, occ = 0, dist_r = 0, range_r = 0, occ_r = 0;
        const bool ok = gp.ComputeOcc(pos, dist, range_pred, occ);
        const bool ok_r = ref.ComputeOcc(pos[0], pos[1], dist_r, range_r, occ_r);
        CHECK(ok == ok_r, "ComputeOcc validity differs");
        if (ok && ok_r) {
            ++occ_checked;
            CHECK(std::abs(double(range_pred) - double(range_r)) < 1e3 * tol * std::max(1.0, std::abs(double(range_r))), "range_pred %g vs %g", double(range_pred), double(range_r));
            CHECK(std::abs(double(occ) - double(occ_r)) < (sizeof(Dtype) == 4 ? 5e-3 : 1e-7), "occ %g vs %g", double(occ), double(occ_r));
        }
    }
    CHECK(occ_checked > 10, "too few valid ComputeOcc samples");
    // GetGps(): the per-partition views a caller of the reference walks (trained flag, n, L, alpha) against the oracle's partition GPs
    const auto &gps = gp.GetGps();
    CHECK(gps.size() == ref.gps.size() && gps.size() == gp.GetAnglePartitions().size(), "GetGps size");
    const double gtol = sizeof(Dtype) == 4 ? 5e-3 : 1e-8;  // alpha carries cond(K)
    for (std::size_t pi = 0; pi < gps.size(); ++pi) {
        const auto &rg = ref.gps[pi];
        CHECK(gps[pi]->IsTrained() == rg.trained, "GetGps()[%zu]->IsTrained", pi);
        if (!rg.trained) { continue; }
        const long np = rg.num_samples;
        CHECK(gps[pi]->GetNumTrainSamples() == np, "GetGps()[%zu] samples %ld vs %ld", pi, gps[pi]->GetNumTrainSamples(), np);
        double el = 0, ea = 0, sa = 0;
        for (long c = 0; c < np; ++c) {
            for (long r = c; r < np; ++r) { el = std::max(el, std::abs(double(gps[pi]->GetCholeskyDecomposition()(r, c)) - double(rg.mat_l[r + c * rg.ld]))); }
            sa = std::max(sa, std::abs(double(rg.mat_alpha[c])));
            ea = std::max(ea, std::abs(double(gps[pi]->GetAlpha()[c]) - double(rg.mat_alpha[c])));
        }
        CHECK(el < (sizeof(Dtype) == 4 ? 2e-5 : 1e-11) && ea / sa < gtol, "GetGps()[%zu]: L err %.3e alpha err %.3e", pi, el, ea / sa);
    }
    std::printf("%s %s: mean err %.2e, var err %.2e, %ld invalid rays\n", g_failures ? "----" : "PASS", name, em / scale, ev, invalid);
}

template<typename Dtype>
static void
TestRangeSensor(const char *name, const double tol) {
    using Rs = RangeSensorGaussianProcess3D<Dtype>;
    constexpr long rows = 72, cols = 40;
    auto setting = std::make_shared<typename Rs::Setting>();
    setting->sensor_frame->azimuth_min = Dtype(-0.5), setting->sensor_frame->azimuth_max = Dtype(0.5), setting->sensor_frame->num_azimuth_lines = rows;
    setting->sensor_frame->elevation_min = Dtype(-0.3), setting->sensor_frame->elevation_max = Dtype(0.3), setting->sensor_frame->num_elevation_lines = cols;
    setting->sensor_frame->valid_range_min = Dtype(0.1), setting->sensor_frame->valid_range_max = Dtype(30);
    setting->gp->kernel_type = "erl::covariance::Matern32<float, 2>";
    setting->gp->kernel->scale = Dtype(0.05);
    Rs gp(setting);
    std::mt19937 rng(5);
    std::uniform_real_distribution<double> uni(0, 1);
    Eigen::MatrixX<Dtype> ranges(rows, cols);
    for (long c = 0; c < cols; ++c) {
        for (long r = 0; r < rows; ++r) { ranges(r, c) = uni(rng) < 0.05 ? Dtype(1e3) : Dtype(4 + 0.8 * std::sin(r / 9.0) * std::cos(c / 13.0)); }
    }
    Eigen::MatrixX<Dtype> rot(3, 3);
    rot.setZero();
    rot(0, 0) = rot(1, 1) = rot(2, 2) = 1;
    Eigen::VectorX<Dtype> trans(3);
    trans.setZero();
    CHECK(gp.Train(rot, trans, ranges), "Train");
    const auto frame = gp.GetSensorFrame();
    erl_gp_oracle::RangeSensorGp3D<Dtype> ref;
    ref.kernel_type = erl_gp_oracle::kMatern32, ref.kernel_scale = Dtype(0.05);
    ref.sensor_range_var = setting->sensor_range_var;
    CHECK(ref.Init(frame->GetFrameCoordsData(), rows, cols), "oracle Init");
    CHECK(gp.GetRowPartitions().size() == ref.row_partitions.size() && gp.GetColPartitions().size() == ref.col_partitions.size(), "partition grid");
    ref.Train(frame->GetRanges().data(), erl::gaussian_process::b200::MaskData(frame->GetHitMask()), true);
    constexpr long n_test = 5000;
    Eigen::MatrixX<Dtype> dirs(3, n_test), coords(2, n_test);
    for (long i = 0; i < n_test; ++i) {
        const double az = -0.55 + 1.1 * uni(rng), el = -0.33 + 0.66 * uni(rng);
        dirs(0, i) = Dtype(std::cos(el) * std::cos(az)), dirs(1, i) = Dtype(std::cos(el) * std::sin(az)), dirs(2, i) = Dtype(std::sin(el));
        Dtype dist;
        (void) frame->ComputeFrameCoords(&dirs(0, i), dist, &coords(0, i));
    }
    auto result = gp.Test(dirs, true, true);
    CHECK(result != nullptr, "Test");
    Eigen::VectorX<Dtype> mean(n_test), var(n_test);
    const auto ok = result->GetMean(mean, true);
    (void) result->GetVariance(var, true);
    std::vector<Dtype> m_ref(n_test), v_ref(n_test);
    std::vector<uint8_t> ok_ref(n_test);
    ref.TestFrameCoords(coords.data(), nullptr, n_test, true, m_ref.data(), v_ref.data(), ok_ref.data());
    double em = 0, ev = 0, scale = 0;
    long valid = 0;
    for (long i = 0; i < n_test; ++i) {
        CHECK(bool(ok[i]) == bool(ok_ref[i]), "valid mask differs at %ld", i);
        if (!ok_ref[i]) { continue; }
        ++valid;
        scale = std::max(scale, std::abs(double(m_ref[i])));
        em = std::max(em, std::abs(double(mean[i]) - double(m_ref[i])));
        ev = std::max(ev, std::abs(double(var[i]) - double(v_ref[i])));
    }
    CHECK(valid > 1000, "too few valid rays: %ld", valid);
    CHECK(em / scale < tol && ev < tol, "mean err %.3e var err %.3e", em / scale, ev);
    // GetGps(): the grid of per-partition views (include/erl_gaussian_process/range_sensor_gp_3d.hpp:136)
    const auto &grid = gp.GetGps();
    CHECK(grid.rows() == long(ref.row_partitions.size()) && grid.cols() == long(ref.col_partitions.size()), "GetGps grid");
    long trained_parts = 0;
    for (long c = 0; c < grid.cols(); c += 3) {
        for (long r = 0; r < grid.rows(); r += 2) {
            const auto &rg = ref.gps[r + c * grid.rows()];
            CHECK(grid(r, c)->IsTrained() == rg.trained, "GetGps()(%ld,%ld)->IsTrained", r, c);
            if (!rg.trained) { continue; }
            ++trained_parts;
            const long np = rg.num_samples;
            CHECK(grid(r, c)->GetNumTrainSamples() == np, "GetGps()(%ld,%ld) samples", r, c);
            double el = 0;
            for (long cc = 0; cc < np; ++cc) {
                for (long rr = cc; rr < np; ++rr) { el = std::max(el, std::abs(double(grid(r, c)->GetCholeskyDecomposition()(rr, cc)) - double(rg.mat_l[rr + cc * rg.ld]))); }
            }
            CHECK(el < (sizeof(Dtype) == 4 ? 5e-5 : 1e-10), "GetGps()(%ld,%ld): L err %.3e", r, c, el);
        }
    }
    CHECK(trained_parts > 3, "too few trained partitions checked");
    std::printf("%s %s: mean err %.2e, var err %.2e, %ld valid rays\n", g_failures ? "----" : "PASS", name, em / scale, ev, valid);
}

template<typename Dtype>
static void
TestSpGpOccupancyMap(const char *name, const double tol) {
    // the flow of src/spgp_occupancy_map.cpp:82-152 on a synthetic 2-D scan: a sensor inside a circular room; two scans from
    // two positions (Q_M and alpha accumulate over Update() calls, src/sparse_pseudo_input_gp.cpp:751-791), then log-odds and
    // their gradient on a grid, against the oracle SPGP fed with the same dataset
    using Map = SpGpOccupancyMap<Dtype, 2>;
    auto setting = std::make_shared<typename Map::Setting>();
    setting->sp_gp->kernel_type = "erl::covariance::Matern32<Dtype, 2>";
    setting->sp_gp->kernel->x_dim = 2;
    setting->sp_gp->kernel->scale = Dtype(1.2);
    setting->sp_gp->max_num_samples = 900;
    setting->min_distance = Dtype(0.3);
    setting->max_distance = Dtype(10);
    setting->logodd_variance = Dtype(0.01);
    constexpr long grid = 12, m = grid * grid;
    Eigen::MatrixX<Dtype> pseudo(2, m);
    for (long i = 0; i < grid; ++i) {
        for (long j = 0; j < grid; ++j) {
            pseudo(0, i * grid + j) = Dtype(-4.4 + 0.8 * i);
            pseudo(1, i * grid + j) = Dtype(-4.4 + 0.8 * j);
        }
    }
    typename Map::AabbD boundary;
    boundary.center.resize(2);
    boundary.half_sizes.resize(2);
    boundary.center[0] = boundary.center[1] = 0;
    boundary.half_sizes[0] = boundary.half_sizes[1] = Dtype(4.5);
    Map map(setting, pseudo, boundary, 7);

    erl_gp_oracle::Spgp<Dtype> ref;
    ref.kernel_type = erl_gp_oracle::kMatern32;
    ref.scale = Dtype(1.2);
    ref.Init(pseudo.data(), 2, m);
    // the same restatement in double on the same (Dtype-rounded) dataset: its distance to `ref` is the floor two correct
    // implementations in Dtype are apart by (M = 144 pseudo-points, noise 0.01: cond(Q_M) ~ 1e5)
    erl_gp_oracle::Spgp<double> ref64;
    ref64.kernel_type = erl_gp_oracle::kMatern32;
    ref64.scale = double(Dtype(1.2));
    {
        std::vector<double> z(static_cast<std::size_t>(2 * m));
        for (long i = 0; i < 2 * m; ++i) { z[i] = pseudo.data()[i]; }
        ref64.Init(z.data(), 2, m);
    }

    Eigen::MatrixX<Dtype> grad_before;
    Eigen::VectorX<Dtype> logodd_before;
    bool threw = false;
    try {
        map.Predict(pseudo, false, true, logodd_before, grad_before);
    } catch (const std::logic_error &) { threw = true; }
    CHECK(threw, "Predict before Update must fail (Test() returns nullptr until trained)");

    const Dtype sensors[2][2] = {{Dtype(0.5), Dtype(-0.3)}, {Dtype(-1.0), Dtype(1.2)}};
    long total_hits = 0;
    for (int scan = 0; scan < 2; ++scan) {
        constexpr long rays = 90;
        Eigen::VectorX<Dtype> sensor(2);
        sensor[0] = sensors[scan][0], sensor[1] = sensors[scan][1];
        Eigen::MatrixX<Dtype> points(2, rays);
        for (long i = 0; i < rays; ++i) {  // hits on the circle |p| = 3.5 (ray / circle intersection)
            const double a = 2 * M_PI * i / rays, dx = std::cos(a), dy = std::sin(a);
            const double b = sensor[0] * dx + sensor[1] * dy, c = sensor[0] * sensor[0] + sensor[1] * sensor[1] - 3.5 * 3.5;
            const double t = -b + std::sqrt(b * b - c);
            points(0, i) = Dtype(sensor[0] + t * dx);
            points(1, i) = Dtype(sensor[1] + t * dy);
        }
        long num_samples = 0;
        Eigen::MatrixX<Dtype> dataset_points;
        Eigen::VectorX<Dtype> dataset_labels;
        std::vector<long> hit_indices;
        CHECK(map.Update(sensor, points, {}, num_samples, dataset_points, dataset_labels, hit_indices), "Update");
        CHECK(num_samples > rays && num_samples <= 900, "dataset size %ld", num_samples);
        CHECK(static_cast<long>(hit_indices.size()) == rays, "%zu hit points", hit_indices.size());
        total_hits += static_cast<long>(hit_indices.size());
        long free_points = 0;
        for (long i = 0; i < num_samples; ++i) {
            const double r = std::hypot(double(dataset_points(0, i)), double(dataset_points(1, i)));
            if (dataset_labels[i] > 0) {
                CHECK(std::abs(r - 3.5) < 1e-3, "hit point off the wall: %f", r);
            } else {
                ++free_points;
                CHECK(r < 3.5, "free point outside the room: %f", r);
            }
        }
        CHECK(free_points > 2 * rays, "%ld free points", free_points);
        std::vector<Dtype> y(static_cast<std::size_t>(num_samples)), var(static_cast<std::size_t>(num_samples), setting->logodd_variance);
        std::vector<Dtype> x(static_cast<std::size_t>(2 * num_samples));
        for (long i = 0; i < num_samples; ++i) {
            y[i] = dataset_labels[i] > 0 ? setting->logodd_occupied : setting->logodd_free;
            x[2 * i] = dataset_points(0, i), x[2 * i + 1] = dataset_points(1, i);
        }
        CHECK(ref.Update(x.data(), y.data(), var.data(), num_samples), "oracle Update");
        std::vector<double> x64(x.begin(), x.end()), y64(y.begin(), y.end()), var64(var.begin(), var.end());
        CHECK(ref64.Update(x64.data(), y64.data(), var64.data(), num_samples), "oracle Update (double)");
    }
    CHECK(map.GetSpGp().IsTrained(), "IsTrained");

    constexpr long tg = 40, nt = tg * tg;
    Eigen::MatrixX<Dtype> xt(2, nt);
    for (long i = 0; i < tg; ++i) {
        for (long j = 0; j < tg; ++j) {
            xt(0, i * tg + j) = Dtype(-4.0 + 8.0 * i / (tg - 1));
            xt(1, i * tg + j) = Dtype(-4.0 + 8.0 * j / (tg - 1));
        }
    }
    Eigen::VectorX<Dtype> logodd;
    Eigen::MatrixX<Dtype> gradient;
    map.Predict(xt, true, true, logodd, gradient);
    std::vector<Dtype> mean_ref(nt), var_ref(nt), grad_ref(2 * nt);
    ref.Test(xt.data(), nt, mean_ref.data(), var_ref.data());
    ref.TestGradient(xt.data(), nt, grad_ref.data(), false);
    std::vector<double> xt64(static_cast<std::size_t>(2 * nt)), mean64(nt), var64(nt), grad64(2 * nt);
    for (long i = 0; i < 2 * nt; ++i) { xt64[i] = xt.data()[i]; }
    ref64.Test(xt64.data(), nt, mean64.data(), var64.data());
    ref64.TestGradient(xt64.data(), nt, grad64.data(), false);
    double em = 0, eg = 0, sm = 0, sg = 0, fm = 0, fg = 0;
    for (long i = 0; i < nt; ++i) {
        em = std::max(em, std::abs(double(logodd[i]) - double(mean_ref[i])));
        fm = std::max(fm, std::abs(mean64[i] - double(mean_ref[i])));
        sm = std::max(sm, std::abs(double(mean_ref[i])));
        for (int d = 0; d < 2; ++d) {
            eg = std::max(eg, std::abs(double(gradient(d, i)) - double(grad_ref[2 * i + d])));
            fg = std::max(fg, std::abs(grad64[2 * i + d] - double(grad_ref[2 * i + d])));
            sg = std::max(sg, std::abs(double(grad_ref[2 * i + d])));
        }
    }
    // bar: the north-star tolerance relative to the field's scale, or 10x the measured floor of the precision
    CHECK(em <= std::max(tol * std::max(1.0, sm), 10 * fm), "log-odds err %.3e (scale %.2f, floor %.3e)", em, sm, fm);
    CHECK(eg <= std::max(tol * std::max(1.0, sg), 10 * fg), "gradient err %.3e (scale %.2f, floor %.3e)", eg, sg, fg);
    // the field separates the two classes: free space near the sensors, occupied on the wall
    Dtype lo_free = 0, lo_wall = 0;
    Eigen::VectorX<Dtype> p(2), g(2);
    p[0] = Dtype(0.5), p[1] = Dtype(-0.3);
    map.Predict(p, true, lo_free, g);
    p[0] = Dtype(3.5), p[1] = Dtype(0);
    map.Predict(p, true, lo_wall, g);
    CHECK(lo_free < -2 && lo_wall > 1, "log-odds at the sensor %.2f, on the wall %.2f", double(lo_free), double(lo_wall));
    CHECK(g[0] > 0, "the log-odds must grow outwards through the wall: d/dx = %.2f", double(g[0]));
    Eigen::MatrixX<Dtype> gradient_only;
    map.PredictGradient(xt, true, gradient_only);
    CHECK(gradient_only == gradient, "PredictGradient must repeat Predict's gradient");
    // variance of the SPGP behind the map and its host-side state
    auto result = map.GetSpGp().Test(xt, false);
    Eigen::VectorX<Dtype> var(nt);
    result->GetVariance(var, true);
    double ev = 0;
    for (long i = 0; i < nt; ++i) { ev = std::max(ev, std::abs(double(var[i]) - double(var_ref[i]))); }
    double fv = 0;
    for (long i = 0; i < nt; ++i) { fv = std::max(fv, std::abs(var64[i] - double(var_ref[i]))); }
    CHECK(ev <= std::max(tol, 10 * fv), "variance err %.3e (floor %.3e)", ev, fv);
    const auto &alpha = map.GetSpGp().GetMatAlpha();
    double ea = 0, sa = 0;
    for (long i = 0; i < m; ++i) {
        ea = std::max(ea, std::abs(double(alpha(i, 0)) - double(ref.alpha[i])));
        sa = std::max(sa, std::abs(double(ref.alpha[i])));
    }
    CHECK(ea <= tol * std::max(1.0, sa), "alpha err %.3e (scale %.2f)", ea, sa);
    // Write / Read / operator== (src/spgp_occupancy_map.cpp:164-254, src/sparse_pseudo_input_gp.cpp:498-749): a map built on other
    // pseudo-points and another seed becomes equal after Read, predicts the same bits, and - the generator travels too - produces the
    // same dataset and the same state from the next scan
    {
        std::stringstream stream;
        CHECK(map.Write(stream), "Write");
        auto setting2 = std::make_shared<typename Map::Setting>();
        setting2->sp_gp->kernel_type = "erl::covariance::RadialBiasFunction<Dtype, 2>";
        setting2->sp_gp->kernel->x_dim = 2;
        Eigen::MatrixX<Dtype> pseudo2(2, 4);
        for (long i = 0; i < 4; ++i) { pseudo2(0, i) = Dtype(i & 1), pseudo2(1, i) = Dtype(i >> 1); }
        Map map2(setting2, pseudo2, boundary, 99);
        CHECK(map != map2, "different maps must differ");
        CHECK(map2.Read(stream), "Read");
        CHECK(map == map2, "Read must restore an equal map");
        CHECK(map2.GetSpGp().IsTrained() && map2.GetSpGp().GetPseudoPoints().cols() == m, "restored SPGP state");
        Eigen::VectorX<Dtype> logodd2;
        Eigen::MatrixX<Dtype> gradient2;
        map2.Predict(xt, true, true, logodd2, gradient2);
        bool same = gradient2 == gradient;
        for (long i = 0; i < nt; ++i) { same = same && logodd2[i] == logodd[i]; }
        CHECK(same, "the restored map must predict the same bits");
        Eigen::VectorX<Dtype> sensor(2);
        sensor[0] = Dtype(1.5), sensor[1] = Dtype(1.0);
        Eigen::MatrixX<Dtype> points(2, 40);
        for (long i = 0; i < 40; ++i) {
            const double a = 2 * M_PI * i / 40;
            points(0, i) = Dtype(3.5 * std::cos(a)), points(1, i) = Dtype(3.5 * std::sin(a));
        }
        long n1 = 0, n2 = 0;
        Eigen::MatrixX<Dtype> dp1, dp2;
        Eigen::VectorX<Dtype> dl1, dl2;
        std::vector<long> h1, h2;
        CHECK(map.Update(sensor, points, {}, n1, dp1, dl1, h1) && map2.Update(sensor, points, {}, n2, dp2, dl2, h2), "Update after Read");
        CHECK(n1 == n2 && h1 == h2 && b200::serialization::SameTopLeft(dp1, dp2, 2, n1), "the restored generator must continue the same stream (%ld vs %ld samples)", n1, n2);
        CHECK(map == map2, "equal maps stay equal under the same update");
        std::stringstream damaged(stream.str().substr(0, stream.str().size() / 2));
        CHECK(!map2.Read(damaged), "a truncated stream must be rejected");
    }
    std::printf("%s %s: log-odds err %.2e (scale %.1f, floor %.1e), gradient err %.2e (scale %.1f, floor %.1e), var err %.2e (floor %.1e), %ld hit points\n", g_failures ? "----" : "PASS", name, em, sm,
                fm, eg, sg, fg, ev, fv, total_hits);
}

int
main() {
    std::setvbuf(stdout, nullptr, _IOLBF, 0);  // progress survives a crash
    try {
        TestVanillaSiso<double>("VanillaGaussianProcess<double> SISO", 1e-10);
        TestVanillaSiso<float>("VanillaGaussianProcess<float> SISO", 1e-4);
        TestLidar<double>("LidarGaussianProcess2D<double>", 1e-10);
        TestLidar<float>("LidarGaussianProcess2D<float>", 1e-4);
        TestNoisyInput<double>("NoisyInputGaussianProcess<double>", 1e-10);
        TestNoisyInput<float>("NoisyInputGaussianProcess<float>", 1e-4);
        TestSerialization<double>("Write / Read / operator== <double>");
        TestSerialization<float>("Write / Read / operator== <float>");
        TestSensorSerialization<double>("sensor GPs Write / Read / operator== <double>");
        TestSensorSerialization<float>("sensor GPs Write / Read / operator== <float>");
        TestLidarHitRays<double>("LidarGaussianProcess2D<double> partition_on_hit_rays", 1e-10);
        TestLidarHitRays<float>("LidarGaussianProcess2D<float> partition_on_hit_rays", 1e-4);
        TestRangeSensor<float>("RangeSensorGaussianProcess3D<float>", 1e-4);
        TestRangeSensor<double>("RangeSensorGaussianProcess3D<double>", 1e-10);
        TestSpGpOccupancyMap<double>("SpGpOccupancyMap<double, 2>", 1e-10);
        TestSpGpOccupancyMap<float>("SpGpOccupancyMap<float, 2>", 1e-4);
        // misuse: hard assertion as ERL_ASSERTM (src/vanilla_gp.cpp:389-392)
        bool threw = false;
        try {
            auto s = std::make_shared<VanillaGaussianProcess<double>::Setting>();
            s->kernel_type = "erl::covariance::Matern32<double, 2>";
            s->max_num_samples = 8;
            VanillaGaussianProcess<double> gp(s);
            gp.Reset(9, 2, 1);
        } catch (const std::logic_error &) { threw = true; }
        CHECK(threw, "Reset beyond max_num_samples must assert");
    } catch (const std::exception &e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf(g_failures == 0 ? "ALL PASS\n" : "%d FAILURES\n", g_failures);
    return g_failures == 0 ? 0 : 1;
}
