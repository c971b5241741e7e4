// VanillaGaussianProcess<Dtype> — drop-in host class over the C ABI.
//
// Same Setting / Reset / GetTrainSet / Train / Test / TestResult::GetMean / GetVariance surface and the same
// state machine as include/erl_gaussian_process/vanilla_gp.hpp + src/vanilla_gp.cpp; the numerics
// (ComputeKtrain, LLT, the two triangular solves, ComputeKtest, mean, variance) run on the GPU through
// erl_gp_vanilla_*.  K, L and alpha stay device resident and are materialised into the host matrices the
// reference exposes (GetKtrain / GetCholeskyDecomposition / GetAlpha) on first access.
#pragma once

#include "c_api.hpp"
#include "covariance.hpp"
#include "eigen_shim.hpp"
#include "serialization.hpp"

#include <memory>
#include <string>
#include <utility>

namespace erl::gaussian_process {

    template<typename Dtype>
    class VanillaGaussianProcess {
    public:
        using Covariance = covariance::Covariance<Dtype>;
        using MatrixX = Eigen::MatrixX<Dtype>;
        using VectorX = Eigen::VectorX<Dtype>;
        using Api = b200::Api<Dtype>;

        struct Setting {
            std::string kernel_type = "erl::covariance::Covariance";
            std::string kernel_setting_type = "erl::covariance::Covariance::Setting";
            std::shared_ptr<typename Covariance::Setting> kernel = std::make_shared<typename Covariance::Setting>();
            long max_num_samples = 256;
        };

        struct TrainSet {
            long x_dim = 0;
            long y_dim = 0;
            long num_samples = 0;
            MatrixX x;    // x_dim x max_num_samples
            MatrixX y;    // max_num_samples x y_dim
            VectorX var;  // max_num_samples

            void
            Reset(const long max_num_samples, const long x_dim_in, const long y_dim_in) {  // grow-only, src/vanilla_gp.cpp:152-161
                x_dim = x_dim_in;
                y_dim = y_dim_in;
                if (x.rows() < x_dim || x.cols() < max_num_samples) { x.resize(x_dim, max_num_samples); }
                if (y.rows() < max_num_samples || y.cols() < y_dim) { y.resize(max_num_samples, y_dim); }
                if (var.size() < max_num_samples) { var.resize(max_num_samples); }
                num_samples = 0;
            }

            [[nodiscard]] bool
            operator==(const TrainSet &other) const {  // src/vanilla_gp.cpp:163-183
                namespace ser = b200::serialization;
                if (x_dim != other.x_dim || y_dim != other.y_dim || num_samples != other.num_samples) { return false; }
                if (num_samples == 0) { return true; }
                if (!ser::SameTopLeft(x, other.x, x_dim, num_samples) || !ser::SameTopLeft(y, other.y, num_samples, y_dim)) { return false; }
                if (var.size() < num_samples || other.var.size() < num_samples) { return false; }
                for (long i = 0; i < num_samples; ++i) {
                    if (!(var[i] == other.var[i])) { return false; }
                }
                return true;
            }

            [[nodiscard]] bool
            operator!=(const TrainSet &other) const {
                return !(*this == other);
            }

            [[nodiscard]] bool
            Write(std::ostream &s) const {  // src/vanilla_gp.cpp:191-235
                namespace ser = b200::serialization;
                return ser::WriteTokens(s, {{"x_dim", ser::ScalarWriter(x_dim)},
                                            {"y_dim", ser::ScalarWriter(y_dim)},
                                            {"num_samples", ser::ScalarWriter(num_samples)},
                                            {"x", [this](std::ostream &o) { return ser::SaveMatrix(o, x); }},
                                            {"y", [this](std::ostream &o) { return ser::SaveMatrix(o, y); }},
                                            {"var", [this](std::ostream &o) { return ser::SaveMatrix(o, var); }}});
            }

            [[nodiscard]] bool
            Read(std::istream &s) {  // src/vanilla_gp.cpp:237-285
                namespace ser = b200::serialization;
                return ser::ReadTokens(s, {{"x_dim", ser::ScalarReader(x_dim)},
                                           {"y_dim", ser::ScalarReader(y_dim)},
                                           {"num_samples", ser::ScalarReader(num_samples)},
                                           {"x", [this](std::istream &i) { return ser::LoadMatrix(i, x); }},
                                           {"y", [this](std::istream &i) { return ser::LoadMatrix(i, y); }},
                                           {"var", [this](std::istream &i) { return ser::LoadVector(i, var); }}});
            }
        };

        class TestResult {
        protected:
            const VanillaGaussianProcess *m_gp_;
            const long m_num_test_;
            const long m_y_dim_;
            MatrixX m_x_test_;
            mutable MatrixX m_mean_;    // num_test x y_dim, lazily computed
            mutable VectorX m_var_;     // num_test, lazily computed (the reference caches L^-1 Kt the same way)

        public:
            TestResult(const VanillaGaussianProcess *gp, const Eigen::Ref<const MatrixX> &mat_x_test)
                : m_gp_(gp),
                  m_num_test_(mat_x_test.cols()),
                  m_y_dim_(gp->m_train_set_.y_dim),
                  m_x_test_(mat_x_test) {}

            [[nodiscard]] long
            GetNumTest() const {
                return m_num_test_;
            }

            void
            GetMean(const long y_index, Eigen::Ref<VectorX> vec_f_out, const bool parallel) const {
                (void) parallel;
                if (m_mean_.size() == 0) {
                    m_mean_.resize(m_num_test_, m_y_dim_);
                    m_gp_->m_ctx_->Check(Api::vanilla_test(m_gp_->m_handle_, m_num_test_, m_x_test_.data(), m_x_test_.rows(), m_mean_.data(), nullptr), "erl_gp_vanilla_test");
                }
                for (long i = 0; i < m_num_test_; ++i) { vec_f_out[i] = m_mean_(i, y_index); }
            }

            void
            GetMean(const long index, const long y_index, Dtype &f) const {
                VectorX tmp(m_num_test_);
                GetMean(y_index, tmp, true);
                f = tmp[index];
            }

            void
            GetVariance(Eigen::Ref<VectorX> vec_var_out, const bool parallel) const {
                (void) parallel;
                if (m_var_.size() == 0) {
                    m_var_.resize(m_num_test_);
                    m_gp_->m_ctx_->Check(Api::vanilla_test(m_gp_->m_handle_, m_num_test_, m_x_test_.data(), m_x_test_.rows(), nullptr, m_var_.data()), "erl_gp_vanilla_test");
                }
                for (long i = 0; i < m_num_test_; ++i) { vec_var_out[i] = m_var_[i]; }
            }

            void
            GetVariance(const long index, Dtype &var) const {
                VectorX tmp(m_num_test_);
                GetVariance(tmp, true);
                var = tmp[index];
            }
        };

    protected:
        std::shared_ptr<Setting> m_setting_ = nullptr;
        std::shared_ptr<b200::DeviceContext> m_ctx_ = nullptr;
        typename Api::Vanilla *m_handle_ = nullptr;
        bool m_trained_ = false;
        bool m_trained_once_ = false;
        bool m_k_train_updated_ = false;
        long m_k_train_rows_ = 0;
        long m_k_train_cols_ = 0;
        int m_llt_info_ = 0;
        mutable bool m_host_copy_valid_ = false;
        mutable MatrixX m_mat_k_train_, m_mat_l_, m_mat_alpha_;
        TrainSet m_train_set_;

    public:
        explicit VanillaGaussianProcess(std::shared_ptr<Setting> setting, std::shared_ptr<b200::DeviceContext> ctx = nullptr)
            : m_setting_(std::move(setting)),
              m_ctx_(ctx ? std::move(ctx) : b200::DeviceContext::Default()) {
            b200::AssertM(m_setting_ != nullptr, "setting should not be nullptr.");
            b200::AssertM(m_setting_->kernel != nullptr, "setting->kernel should not be nullptr.");
            m_ctx_->Check(Api::vanilla_create(m_ctx_->Get(), &m_handle_), "erl_gp_vanilla_create");
        }

        VanillaGaussianProcess(const VanillaGaussianProcess &) = delete;
        VanillaGaussianProcess &
        operator=(const VanillaGaussianProcess &) = delete;

        ~VanillaGaussianProcess() { Api::vanilla_destroy(m_handle_); }

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
            return false;  // the three north-star kernels are full rank (src/vanilla_gp.cpp:825-828)
        }

        void
        Reset(const long max_num_samples, const long x_dim, const long y_dim) {  // src/vanilla_gp.cpp:376-400
            b200::AssertM(max_num_samples > 0, "max_num_samples should be > 0.");
            b200::AssertM(x_dim > 0, "x_dim should be > 0.");
            b200::AssertM(y_dim > 0, "y_dim should be > 0.");
            b200::AssertM(m_setting_->kernel->x_dim == -1 || m_setting_->kernel->x_dim == x_dim, "x_dim should be " + std::to_string(m_setting_->kernel->x_dim) + ".");
            b200::AssertM(m_setting_->max_num_samples < 0 || max_num_samples <= m_setting_->max_num_samples,
                          "max_num_samples should be <= " + std::to_string(m_setting_->max_num_samples) + ".");
            m_train_set_.Reset(max_num_samples, x_dim, y_dim);
            m_trained_ = false;
            m_k_train_updated_ = false;
            m_k_train_rows_ = 0;
            m_k_train_cols_ = 0;
            m_host_copy_valid_ = false;
        }

        [[nodiscard]] std::pair<long, long>
        GetKtrainSize() const {
            return {m_k_train_rows_, m_k_train_cols_};
        }

        [[nodiscard]] TrainSet &
        GetTrainSet() {
            return m_train_set_;
        }

        [[nodiscard]] const TrainSet &
        GetTrainSet() const {
            return m_train_set_;
        }

        [[nodiscard]] const MatrixX &
        GetKtrain() const {
            Materialise();
            return m_mat_k_train_;
        }

        [[nodiscard]] const MatrixX &
        GetCholeskyDecomposition() const {
            Materialise();
            return m_mat_l_;
        }

        [[nodiscard]] const MatrixX &
        GetAlpha() const {
            Materialise();
            return m_mat_alpha_;
        }

        // info of the factorisation: 0, or k > 0 if the leading minor of order k is not positive.  The reference
        // ignores Eigen's llt().info() (src/vanilla_gp.cpp:499); this accessor is an addition, Train() still returns true.
        [[nodiscard]] int
        GetLltInfo() const {
            return m_llt_info_;
        }

        [[nodiscard]] bool
        Train() {  // src/vanilla_gp.cpp:507-519 with UpdateKtrain (:476-490) and Solve (:492-505) fused on the device
            if (m_trained_) { return false; }  // "The model has been trained. Please reset the model before training."
            m_trained_ = m_trained_once_;
            auto &[x_dim, y_dim, num_samples, x, y, var] = m_train_set_;
            if (!m_k_train_updated_) {
                if (num_samples <= 0) { return false; }
                const int kernel = covariance::KernelFromTypeName(m_setting_->kernel_type);
                m_ctx_->Check(
                    Api::vanilla_train(m_handle_, kernel, m_setting_->kernel->scale, x_dim, y_dim, num_samples, x.data(), x.rows(), y.data(), y.rows(), var.data(), &m_llt_info_),
                    "erl_gp_vanilla_train");
                m_k_train_rows_ = m_k_train_cols_ = num_samples;
                m_k_train_updated_ = true;
                m_host_copy_valid_ = false;
            }
            m_trained_once_ = true;
            m_trained_ = true;
            return true;
        }

        [[nodiscard]] std::shared_ptr<TestResult>
        Test(const Eigen::Ref<const MatrixX> &mat_x_test) const {
            if (!m_trained_) { return nullptr; }          // src/vanilla_gp.cpp:556-558
            if (mat_x_test.cols() == 0) { return nullptr; }  // ComputeKtest warns and fails on num_test == 0 (:529-532)
            return std::make_shared<TestResult>(this, mat_x_test);
        }

        // ---- operator== / Write / Read (src/vanilla_gp.cpp:561-790): same tokens in the same order; see serialization.hpp ----
        [[nodiscard]] bool
        operator==(const VanillaGaussianProcess &other) const {
            namespace ser = b200::serialization;
            if (!ser::SameGpSetting(*m_setting_, *other.m_setting_)) { return false; }
            if (m_trained_ != other.m_trained_ || m_trained_once_ != other.m_trained_once_ || m_k_train_updated_ != other.m_k_train_updated_) { return false; }
            if (m_k_train_rows_ != other.m_k_train_rows_ || m_k_train_cols_ != other.m_k_train_cols_) { return false; }
            if (m_train_set_ != other.m_train_set_) { return false; }
            if (!m_k_train_updated_) { return true; }
            Materialise();
            other.Materialise();
            return ser::SameTopLeft(m_mat_k_train_, other.m_mat_k_train_, m_k_train_rows_, m_k_train_cols_) &&
                   ser::SameTopLeft(m_mat_l_, other.m_mat_l_, m_k_train_rows_, m_k_train_cols_) && m_mat_alpha_.cols() == other.m_mat_alpha_.cols() &&
                   ser::SameTopLeft(m_mat_alpha_, other.m_mat_alpha_, m_k_train_cols_, m_mat_alpha_.cols());
        }

        [[nodiscard]] bool
        operator!=(const VanillaGaussianProcess &other) const {
            return !(*this == other);
        }

        [[nodiscard]] bool
        Write(std::ostream &s) const {
            namespace ser = b200::serialization;
            Materialise();
            return ser::WriteTokens(s, {{"setting", [this](std::ostream &o) { return ser::WriteGpSetting(o, *m_setting_); }},
                                        {"trained", ser::ScalarWriter(m_trained_)},
                                        {"trained_once", ser::ScalarWriter(m_trained_once_)},
                                        {"k_train_updated", ser::ScalarWriter(m_k_train_updated_)},
                                        {"k_train_rows", ser::ScalarWriter(m_k_train_rows_)},
                                        {"k_train_cols", ser::ScalarWriter(m_k_train_cols_)},
                                        {"kernel", [](std::ostream &o) { o << true << '\n'; return o.good(); }},  // the kernel's own state is its Setting, written above
                                        {"mat_k_train", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_k_train_); }},
                                        {"mat_l", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_l_); }},
                                        {"mat_alpha", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_alpha_); }},
                                        {"train_set", [this](std::ostream &o) { return m_train_set_.Write(o); }}});
        }

        [[nodiscard]] bool
        Read(std::istream &s) {
            namespace ser = b200::serialization;
            bool trained = false, trained_once = false, updated = false, has_kernel = false;
            long rows = 0, cols = 0;
            MatrixX k, l, alpha;
            const bool ok = ser::ReadTokens(s, {{"setting", [this](std::istream &i) { return ser::ReadGpSetting(i, *m_setting_); }},
                                                {"trained", ser::ScalarReader(trained)},
                                                {"trained_once", ser::ScalarReader(trained_once)},
                                                {"k_train_updated", ser::ScalarReader(updated)},
                                                {"k_train_rows", ser::ScalarReader(rows)},
                                                {"k_train_cols", ser::ScalarReader(cols)},
                                                {"kernel", [&has_kernel](std::istream &i) { i >> has_kernel; ser::SkipLine(i); return !i.fail(); }},
                                                {"mat_k_train", [&k](std::istream &i) { return ser::LoadMatrix(i, k); }},
                                                {"mat_l", [&l](std::istream &i) { return ser::LoadMatrix(i, l); }},
                                                {"mat_alpha", [&alpha](std::istream &i) { return ser::LoadMatrix(i, alpha); }},
                                                {"train_set", [this](std::istream &i) { return m_train_set_.Read(i); }}});
            if (!ok) { return false; }
            m_trained_ = false;
            m_k_train_updated_ = false;
            m_host_copy_valid_ = false;
            if (updated) {  // rebuild the device state: the training is bit-reproducible, the stored L is the check
                m_trained_once_ = false;
                if (!Train()) { return false; }
                Materialise();
                if (m_k_train_rows_ != rows || m_k_train_cols_ != cols || !ser::SameTopLeft(m_mat_l_, l, rows, cols) || !ser::SameTopLeft(m_mat_alpha_, alpha, cols, alpha.cols())) { return false; }
            }
            m_trained_ = trained;
            m_trained_once_ = trained_once;
            m_k_train_updated_ = updated;
            return true;
        }

    protected:
        void
        Materialise() const {
            if (m_host_copy_valid_ || !m_k_train_updated_) { return; }
            const long n = m_k_train_rows_;
            m_mat_k_train_.resize(n, n);
            m_mat_l_.resize(n, n);
            m_mat_alpha_.resize(n, m_train_set_.y_dim);
            m_ctx_->Check(Api::vanilla_get(m_handle_, m_mat_k_train_.data(), n, m_mat_l_.data(), n, m_mat_alpha_.data(), n), "erl_gp_vanilla_get");
            m_host_copy_valid_ = true;
        }
    };

    using VanillaGaussianProcessD = VanillaGaussianProcess<double>;
    using VanillaGaussianProcessF = VanillaGaussianProcess<float>;
}  // namespace erl::gaussian_process
