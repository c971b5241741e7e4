// SparsePseudoInputGaussianProcess<Dtype> — drop-in host class over the C ABI (erl_gp_spgp_*).
//
// Same Setting / TrainSet / Reset / Update / Test / TestResult surface and state machine as
// include/erl_gaussian_process/sparse_pseudo_input_gp.hpp + src/sparse_pseudo_input_gp.cpp, dense mode: K_M and its factor at
// construction (:313-356), every Update() accumulates Q_M and alpha over the new samples (:751-791), L_QM is refactored lazily
// at the first Test() after an update (:835-842).  All of it runs on the GPU; Q_M, alpha, L_KM and L_QM come to the host on first
// access.  Built for y_dim = 1 (what SpGpOccupancyMap uses); `use_sparse` is rejected (ComputeKtestSparse is not part of this
// build); `diagonal_qm` serves mean and gradient, its variance is rejected as in-reference-undefined.
#pragma once

#include "c_api.hpp"
#include "covariance.hpp"
#include "eigen_shim.hpp"
#include "serialization.hpp"
#include "vanilla_gp.hpp"

#include <memory>
#include <string>
#include <type_traits>
#include <utility>

namespace erl::gaussian_process {

    template<typename Dtype>
    class SparsePseudoInputGaussianProcess {
    public:
        using Covariance = covariance::Covariance<Dtype>;
        using MatrixX = Eigen::MatrixX<Dtype>;
        using VectorX = Eigen::VectorX<Dtype>;
        using TrainSet = typename VanillaGaussianProcess<Dtype>::TrainSet;  // sparse_pseudo_input_gp.hpp:43

        struct Setting {  // sparse_pseudo_input_gp.hpp:45-57
            std::string kernel_type = "erl::covariance::Covariance";
            std::string kernel_setting_type = "erl::covariance::Covariance::Setting";
            std::shared_ptr<typename Covariance::Setting> kernel = std::make_shared<typename Covariance::Setting>();
            long max_num_samples = 256;
            Dtype sparse_zero_threshold = 1e-6f;
            bool use_sparse = false;
            bool diagonal_qm = false;
        };

    private:
        // typed entry points of the C ABI (the SPGP block of include/erl_gp_b200.h)
        struct Abi;

    public:
        class TestResult {  // sparse_pseudo_input_gp.hpp:68-117
            const SparsePseudoInputGaussianProcess *m_gp_;
            long m_num_test_;
            bool m_support_gradient_;
            long m_x_dim_;
            MatrixX m_x_test_;  // x_dim x num_test
            mutable VectorX m_mean_, m_var_;
            mutable MatrixX m_grad_;
            mutable bool m_have_mean_ = false, m_have_var_ = false, m_have_grad_ = false;

            void
            FetchMean(const bool with_variance) const {
                if (m_have_mean_ && (m_have_var_ || !with_variance)) { return; }
                m_mean_.resize(m_num_test_);
                if (with_variance) { m_var_.resize(m_num_test_); }
                m_gp_->Check(Abi::test(m_gp_->m_handle_, m_num_test_, m_x_test_.data(), m_x_dim_, m_mean_.data(), with_variance ? m_var_.data() : nullptr), "erl_gp_spgp_test");
                m_have_mean_ = true;
                m_have_var_ = m_have_var_ || with_variance;
            }

            void
            FetchGradient() const {
                if (m_have_grad_) { return; }
                b200::AssertM(m_support_gradient_, "m_support_gradient_ = false, it should be true to call GetGradient().");
                m_grad_.resize(m_x_dim_, m_num_test_);
                // dotted with Q_M^-1 alpha, consistent with GetMean (the per-index accessor of the reference, :252)
                m_gp_->Check(Abi::test_gradient(m_gp_->m_handle_, m_num_test_, m_x_test_.data(), m_x_dim_, m_grad_.data(), 0), "erl_gp_spgp_test_gradient");
                m_have_grad_ = true;
            }

        public:
            TestResult(const SparsePseudoInputGaussianProcess *gp, const Eigen::Ref<const MatrixX> &mat_x_test, const bool will_predict_gradient)
                : m_gp_(gp),
                  m_num_test_(mat_x_test.cols()),
                  m_support_gradient_(will_predict_gradient),
                  m_x_dim_(gp->m_pseudo_points_.rows()),
                  m_x_test_(mat_x_test) {
                b200::AssertM(mat_x_test.rows() == m_x_dim_, "mat_x_test.rows() should be " + std::to_string(m_x_dim_));
            }

            [[nodiscard]] long
            GetNumTest() const {
                return m_num_test_;
            }

            [[nodiscard]] long
            GetDimX() const {
                return m_x_dim_;
            }

            [[nodiscard]] long
            GetDimY() const {
                return 1;
            }

            void
            GetMean(const long y_index, Eigen::Ref<VectorX> vec_f_out, const bool /*parallel*/) const {  // :43-80
                b200::AssertM(y_index == 0, "y_index should be 0 (y_dim = 1)");
                FetchMean(false);
                for (long i = 0; i < m_num_test_; ++i) { vec_f_out[i] = m_mean_[i]; }
            }

            void
            GetMean(const long index, const long y_index, Dtype &f) const {  // :82-113
                b200::AssertM(y_index == 0 && index >= 0 && index < m_num_test_, "index / y_index out of range");
                FetchMean(false);
                f = m_mean_[index];
            }

            // every gradient is valid for the stationary kernels of this build (the reference's flags mark the reduced-rank failures)
            [[nodiscard]] Eigen::VectorXb
            GetGradient(const long y_index, Eigen::Ref<MatrixX> mat_grad_out, const bool /*parallel*/) const {  // :187-234
                b200::AssertM(y_index == 0, "y_index should be 0 (y_dim = 1)");
                FetchGradient();
                Eigen::VectorXb valid_gradients;
                valid_gradients.resize(m_num_test_);
                for (long i = 0; i < m_num_test_; ++i) {
                    for (long d = 0; d < m_x_dim_; ++d) { mat_grad_out(d, i) = m_grad_(d, i); }
                    valid_gradients[i] = 1;
                }
                return valid_gradients;
            }

            [[nodiscard]] bool
            GetGradient(const long index, const long y_index, Dtype *grad) const {  // :236-278
                b200::AssertM(y_index == 0 && index >= 0 && index < m_num_test_, "index / y_index out of range");
                FetchGradient();
                for (long d = 0; d < m_x_dim_; ++d) { grad[d] = m_grad_(d, index); }
                return true;
            }

            void
            GetVariance(Eigen::Ref<VectorX> vec_var_out, const bool /*parallel*/) const {  // :280-293
                FetchMean(true);
                for (long i = 0; i < m_num_test_; ++i) { vec_var_out[i] = m_var_[i]; }
            }

            void
            GetVariance(const long index, Dtype &var) const {  // :295-300
                FetchMean(true);
                var = m_var_[index];
            }
        };

        SparsePseudoInputGaussianProcess() = delete;

        explicit SparsePseudoInputGaussianProcess(std::shared_ptr<Setting> setting, MatrixX pseudo_points, std::shared_ptr<b200::DeviceContext> ctx = nullptr)
            : m_setting_(std::move(setting)),
              m_pseudo_points_(std::move(pseudo_points)),
              m_ctx_(ctx != nullptr ? std::move(ctx) : b200::DeviceContext::Default()) {
            b200::AssertM(m_setting_ != nullptr, "setting is null");
            b200::AssertM(m_setting_->kernel != nullptr, "setting->kernel is null");
            b200::AssertM(m_pseudo_points_.cols() > 0, "pseudo_points must have at least one column");
            b200::AssertM(m_setting_->kernel->x_dim == -1 || m_setting_->kernel->x_dim == m_pseudo_points_.rows(), "setting->kernel->x_dim and pseudo_points.rows() should match.");
            b200::AssertM(!m_setting_->use_sparse, "use_sparse = true is not part of this build (ComputeKtestSparse of erl_covariance is absent)");
            const int kernel = covariance::KernelFromTypeName(m_setting_->kernel_type);
            Check(Abi::create(m_ctx_->Get(), kernel, m_setting_->kernel->scale, m_pseudo_points_.rows(), m_pseudo_points_.cols(), m_pseudo_points_.data(), &m_handle_), "erl_gp_spgp_create");
            if (m_setting_->diagonal_qm) { Check(Abi::set_diagonal_qm(m_handle_, 1), "erl_gp_spgp_set_diagonal_qm"); }
        }

        SparsePseudoInputGaussianProcess(const SparsePseudoInputGaussianProcess &) = delete;
        SparsePseudoInputGaussianProcess &
        operator=(const SparsePseudoInputGaussianProcess &) = delete;

        ~SparsePseudoInputGaussianProcess() {
            if (m_handle_ != nullptr) { Abi::destroy(m_handle_); }
        }

        [[nodiscard]] std::shared_ptr<const Setting>
        GetSetting() const {
            return m_setting_;
        }

        [[nodiscard]] bool
        IsTrained() const {
            return m_trained_;
        }

        [[nodiscard]] bool
        UsingReducedRankKernel() const {
            return false;
        }

        [[nodiscard]] VectorX
        GetKernelCoordOrigin() const {  // :201-210
            VectorX origin(m_pseudo_points_.rows());
            for (long d = 0; d < origin.size(); ++d) { origin[d] = 0; }
            return origin;
        }

        void
        SetKernelCoordOrigin(const VectorX & /*coord_origin*/) const {}  // only reduced-rank kernels have one (:212-220)

        void
        Reset(const long max_num_samples, const long x_dim, const long y_dim) {  // :395-420: the train set only, Q_M and alpha keep accumulating
            b200::AssertM(max_num_samples > 0, "max_num_samples should be > 0.");
            b200::AssertM(x_dim == m_pseudo_points_.rows(), "x_dim should be " + std::to_string(m_pseudo_points_.rows()));
            b200::AssertM(y_dim == 1, "this build serves y_dim = 1");
            b200::AssertM(m_setting_->max_num_samples < 0 || max_num_samples <= m_setting_->max_num_samples, "max_num_samples should be <= " + std::to_string(m_setting_->max_num_samples));
            m_trained_ = false;
            m_train_set_.Reset(max_num_samples, x_dim, y_dim);
        }

        [[nodiscard]] const MatrixX &
        GetPseudoPoints() const {
            return m_pseudo_points_;
        }

        [[nodiscard]] const MatrixX &
        GetMatLKm() const {
            Fetch();
            return m_mat_l_km_;
        }

        [[nodiscard]] const MatrixX &
        GetMatQm() const {  // M x M; M x 1 with diagonal_qm (:346-347)
            Fetch();
            return m_mat_qm_;
        }

        [[nodiscard]] const MatrixX &
        GetMatLQm() const {
            Fetch();
            return m_mat_l_qm_;
        }

        [[nodiscard]] const MatrixX &
        GetMatAlpha() const {
            Fetch();
            return m_mat_alpha_;
        }

        [[nodiscard]] TrainSet &
        GetTrainSet() {
            return m_train_set_;
        }

        [[nodiscard]] const TrainSet &
        GetTrainSet() const {
            return m_train_set_;
        }

        [[nodiscard]] bool
        Update(const bool /*parallel*/) {  // :751-791
            m_trained_ = m_trained_once_;
            if (m_train_set_.num_samples == 0) { return false; }
            const long n = m_train_set_.num_samples;
            Check(Abi::update(m_handle_, n, m_train_set_.x.data(), m_train_set_.x.rows(), m_train_set_.y.data(), m_train_set_.var.data()), "erl_gp_spgp_update");
            m_host_valid_ = false;
            m_trained_once_ = true;
            m_trained_ = true;
            return true;
        }

        [[nodiscard]] std::shared_ptr<TestResult>
        Test(const Eigen::Ref<const MatrixX> &mat_x_test, const bool predict_gradient) const {  // :793-803
            if (!m_trained_) { return nullptr; }
            return std::make_shared<TestResult>(this, mat_x_test, predict_gradient);
        }

        [[nodiscard]] bool
        operator==(const SparsePseudoInputGaussianProcess &other) const {  // :498-532
            namespace ser = b200::serialization;
            if (!SameSetting(*m_setting_, *other.m_setting_)) { return false; }
            if (m_trained_ != other.m_trained_ || m_trained_once_ != other.m_trained_once_) { return false; }
            if (!(m_pseudo_points_ == other.m_pseudo_points_)) { return false; }
            Fetch();
            other.Fetch();
            if (!(m_mat_qm_ == other.m_mat_qm_) || !(m_mat_alpha_ == other.m_mat_alpha_) || !(m_mat_l_km_ == other.m_mat_l_km_)) { return false; }
            return m_train_set_ == other.m_train_set_;
        }

        [[nodiscard]] bool
        operator!=(const SparsePseudoInputGaussianProcess &other) const {
            return !(*this == other);
        }

        // Token stream as src/sparse_pseudo_input_gp.cpp:539-640 (own framing, serialization.hpp).  K_M, L_KM and L_QM are not
        // stored: they follow from the pseudo-points and Q_M, which Read() hands to a fresh device handle.
        [[nodiscard]] bool
        Write(std::ostream &s) const {
            namespace ser = b200::serialization;
            Fetch();
            return ser::WriteTokens(s, {{"setting", [this](std::ostream &o) { return ser::WriteGpSetting(o, *m_setting_); }},
                                        {"spgp_setting",
                                         [this](std::ostream &o) {
                                             o.precision(17);
                                             o << m_setting_->sparse_zero_threshold << ' ' << m_setting_->use_sparse << ' ' << m_setting_->diagonal_qm;
                                             return o.good();
                                         }},
                                        {"trained", ser::ScalarWriter(m_trained_)},
                                        {"trained_once", ser::ScalarWriter(m_trained_once_)},
                                        {"pseudo_points", [this](std::ostream &o) { return ser::SaveMatrix(o, m_pseudo_points_); }},
                                        {"mat_qm", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_qm_); }},
                                        {"mat_alpha", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_alpha_); }},
                                        {"train_set", [this](std::ostream &o) { return m_train_set_.Write(o); }}});
        }

        [[nodiscard]] bool
        Read(std::istream &s) {  // :642-749
            namespace ser = b200::serialization;
            bool trained = false, trained_once = false;
            MatrixX pseudo, qm, alpha;
            auto setting = std::make_shared<Setting>();
            const bool ok = ser::ReadTokens(s, {{"setting", [&setting](std::istream &i) { return ser::ReadGpSetting(i, *setting); }},
                                                {"spgp_setting",
                                                 [&setting](std::istream &i) {
                                                     i >> setting->sparse_zero_threshold >> setting->use_sparse >> setting->diagonal_qm;
                                                     return !i.fail();
                                                 }},
                                                {"trained", ser::ScalarReader(trained)},
                                                {"trained_once", ser::ScalarReader(trained_once)},
                                                {"pseudo_points", [&pseudo](std::istream &i) { return ser::LoadMatrix(i, pseudo); }},
                                                {"mat_qm", [&qm](std::istream &i) { return ser::LoadMatrix(i, qm); }},
                                                {"mat_alpha", [&alpha](std::istream &i) { return ser::LoadMatrix(i, alpha); }},
                                                {"train_set", [this](std::istream &i) { return m_train_set_.Read(i); }}});
            if (!ok || setting->use_sparse) { return false; }
            const long m = pseudo.cols();
            if (m <= 0 || pseudo.rows() < 1 || alpha.rows() != m || alpha.cols() != 1 || qm.rows() != m || qm.cols() != (setting->diagonal_qm ? 1 : m)) { return false; }
            int kernel = 0;
            try {
                kernel = covariance::KernelFromTypeName(setting->kernel_type);
            } catch (const std::logic_error &) { return false; }
            // a fresh device handle for the stored pseudo-points (K_M, L_KM), then the accumulated state
            typename Abi::Handle *handle = nullptr;
            if (Abi::create(m_ctx_->Get(), kernel, setting->kernel->scale, pseudo.rows(), m, pseudo.data(), &handle) != ERL_GP_STATUS_OK) { return false; }
            if ((setting->diagonal_qm && Abi::set_diagonal_qm(handle, 1) != ERL_GP_STATUS_OK) || Abi::set_state(handle, qm.data(), alpha.data()) != ERL_GP_STATUS_OK) {
                Abi::destroy(handle);
                return false;
            }
            if (m_handle_ != nullptr) { Abi::destroy(m_handle_); }
            m_handle_ = handle;
            *m_setting_ = *setting;  // the caller's shared Setting object follows the stream, like Yamlable::Read in the reference
            m_pseudo_points_ = pseudo;
            m_trained_ = trained;
            m_trained_once_ = trained_once;
            m_host_valid_ = false;
            return true;
        }

    private:
        static bool
        SameSetting(const Setting &a, const Setting &b) {
            return b200::serialization::SameGpSetting(a, b) && a.sparse_zero_threshold == b.sparse_zero_threshold && a.use_sparse == b.use_sparse && a.diagonal_qm == b.diagonal_qm;
        }

        struct Abi {
            using Handle = std::conditional_t<std::is_same_v<Dtype, float>, erl_gp_spgp_f32, erl_gp_spgp_f64>;
            template<typename F32, typename F64>
            static constexpr auto
            Pick(F32 f32, F64 f64) {
                if constexpr (std::is_same_v<Dtype, float>) {
                    return f32;
                } else {
                    return f64;
                }
            }
            static constexpr auto create = Pick(erl_gp_spgp_create_f32, erl_gp_spgp_create_f64);
            static constexpr auto destroy = Pick(erl_gp_spgp_destroy_f32, erl_gp_spgp_destroy_f64);
            static constexpr auto update = Pick(erl_gp_spgp_update_f32, erl_gp_spgp_update_f64);
            static constexpr auto test = Pick(erl_gp_spgp_test_f32, erl_gp_spgp_test_f64);
            static constexpr auto test_gradient = Pick(erl_gp_spgp_test_gradient_f32, erl_gp_spgp_test_gradient_f64);
            static constexpr auto set_diagonal_qm = Pick(erl_gp_spgp_set_diagonal_qm_f32, erl_gp_spgp_set_diagonal_qm_f64);
            static constexpr auto get_qm_diagonal = Pick(erl_gp_spgp_get_qm_diagonal_f32, erl_gp_spgp_get_qm_diagonal_f64);
            static constexpr auto get = Pick(erl_gp_spgp_get_f32, erl_gp_spgp_get_f64);
            static constexpr auto set_state = Pick(erl_gp_spgp_set_state_f32, erl_gp_spgp_set_state_f64);
        };

        void
        Check(const int rc, const char *where) const {
            m_ctx_->Check(rc, where);
        }

        void
        Fetch() const {
            if (m_host_valid_) { return; }
            const long m = m_pseudo_points_.cols();
            m_mat_alpha_.resize(m, 1);
            m_mat_l_km_.resize(m, m);
            if (m_setting_->diagonal_qm) {
                m_mat_qm_.resize(m, 1);
                Check(Abi::get_qm_diagonal(m_handle_, m_mat_qm_.data()), "erl_gp_spgp_get_qm_diagonal");
                Check(Abi::get(m_handle_, nullptr, m_mat_alpha_.data(), m_mat_l_km_.data(), nullptr), "erl_gp_spgp_get");
            } else {
                m_mat_qm_.resize(m, m);
                m_mat_l_qm_.resize(m, m);
                Check(Abi::get(m_handle_, m_mat_qm_.data(), m_mat_alpha_.data(), m_mat_l_km_.data(), m_mat_l_qm_.data()), "erl_gp_spgp_get");
            }
            m_host_valid_ = true;
        }

        std::shared_ptr<Setting> m_setting_ = nullptr;
        bool m_trained_ = false;
        bool m_trained_once_ = false;
        MatrixX m_pseudo_points_{};
        std::shared_ptr<b200::DeviceContext> m_ctx_;
        typename Abi::Handle *m_handle_ = nullptr;
        TrainSet m_train_set_;
        mutable bool m_host_valid_ = false;
        mutable MatrixX m_mat_qm_{}, m_mat_l_km_{}, m_mat_l_qm_{}, m_mat_alpha_{};
    };

    using SparsePseudoInputGaussianProcessD = SparsePseudoInputGaussianProcess<double>;
    using SparsePseudoInputGaussianProcessF = SparsePseudoInputGaussianProcess<float>;

}  // namespace erl::gaussian_process
