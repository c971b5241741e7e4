// NoisyInputGaussianProcess<Dtype> — drop-in host class over the C ABI (erl_gp_noisy_*).
//
// Same Setting / TrainSet / Reset / Train / Test / TestResult surface and state machine as
// include/erl_gaussian_process/noisy_input_gp.hpp + src/noisy_input_gp.cpp.  UpdateKtrain (the derivative-augmented Gram
// matrix of erl_covariance::ComputeKtrainWithGradient), the LLT, the two triangular solves, Ktest with gradient columns and
// every TestResult output run on the GPU; K, L and alpha are materialised on the host on first access.
#pragma once

#include "c_api.hpp"
#include "covariance.hpp"
#include "eigen_shim.hpp"
#include "serialization.hpp"

#include <cmath>
#include <memory>
#include <string>
#include <utility>

namespace erl::gaussian_process {

    template<typename Dtype>
    class NoisyInputGaussianProcess {
    public:
        using Covariance = covariance::Covariance<Dtype>;
        using MatrixX = Eigen::MatrixX<Dtype>;
        using VectorX = Eigen::VectorX<Dtype>;
        using VectorXl = Eigen::VectorX<long>;
        using Api = b200::Api<Dtype>;

        struct Setting {  // noisy_input_gp.hpp:20-33
            std::string kernel_type = "erl::covariance::Covariance";
            std::string kernel_setting_type = "erl::covariance::Covariance::Setting";
            std::shared_ptr<typename Covariance::Setting> kernel = std::make_shared<typename Covariance::Setting>();
            long max_num_samples = -1;  // -1 = no limit
            bool no_gradient_observation = false;
        };

        struct TrainSet {  // noisy_input_gp.hpp:166-199
            long x_dim = 0;
            long y_dim = 0;
            long num_samples = 0;
            long num_samples_with_grad = 0;
            MatrixX x;           // x_dim x max_num_samples
            MatrixX y;           // max_num_samples x y_dim
            MatrixX grad;        // (x_dim * y_dim) x max_num_samples: column i = dh_1/dx_1 .. dh_1/dx_m, dh_2/dx_1 ..
            VectorX var_x;       // input noise
            VectorX var_y;       // output noise
            VectorX var_grad;    // gradient noise
            VectorXl grad_flag;  // != 0: the sample has a gradient observation

            void
            Reset(const long max_num_samples, const long x_dim_in, const long y_dim_in, const bool no_gradient_observation) {  // grow-only, src/noisy_input_gp.cpp:379-402
                x_dim = x_dim_in;
                y_dim = y_dim_in;
                if (x.rows() < x_dim || x.cols() < max_num_samples) { x.resize(x_dim, max_num_samples); }
                if (y.rows() < max_num_samples || y.cols() < y_dim) { y.resize(max_num_samples, y_dim); }
                if (grad_flag.size() < max_num_samples) { grad_flag.resize(max_num_samples); }
                if (var_x.size() < max_num_samples) { var_x.resize(max_num_samples); }
                if (var_y.size() < max_num_samples) { var_y.resize(max_num_samples); }
                if (!no_gradient_observation) {
                    if (grad.rows() < x_dim * y_dim || grad.cols() < max_num_samples) { grad.resize(x_dim * y_dim, max_num_samples); }
                    if (var_grad.size() < max_num_samples) { var_grad.resize(max_num_samples); }
                }
                num_samples = 0;
                num_samples_with_grad = 0;
            }

            [[nodiscard]] bool
            operator==(const TrainSet &other) const {  // src/noisy_input_gp.cpp:404-434
                namespace ser = b200::serialization;
                if (x_dim != other.x_dim || y_dim != other.y_dim || num_samples != other.num_samples || num_samples_with_grad != other.num_samples_with_grad) { return false; }
                if (num_samples == 0) { return true; }
                if (!ser::SameTopLeft(x, other.x, x_dim, num_samples) || !ser::SameTopLeft(y, other.y, num_samples, y_dim)) { return false; }
                if (grad_flag.size() < num_samples || other.grad_flag.size() < num_samples) { return false; }
                for (long i = 0; i < num_samples; ++i) {
                    if (grad_flag[i] != other.grad_flag[i] || !(var_x[i] == other.var_x[i]) || !(var_y[i] == other.var_y[i])) { return false; }
                }
                if (num_samples_with_grad == 0) { return true; }
                if (!ser::SameTopLeft(grad, other.grad, x_dim * y_dim, num_samples)) { return false; }
                for (long i = 0; i < num_samples; ++i) {
                    if (!(var_grad[i] == other.var_grad[i])) { return false; }
                }
                return true;
            }

            [[nodiscard]] bool
            operator!=(const TrainSet &other) const {
                return !(*this == other);
            }

            [[nodiscard]] bool
            Write(std::ostream &s) const {  // src/noisy_input_gp.cpp:436-517
                namespace ser = b200::serialization;
                return ser::WriteTokens(s, {{"x_dim", ser::ScalarWriter(x_dim)},
                                            {"y_dim", ser::ScalarWriter(y_dim)},
                                            {"num_samples", ser::ScalarWriter(num_samples)},
                                            {"num_samples_with_grad", ser::ScalarWriter(num_samples_with_grad)},
                                            {"x", [this](std::ostream &o) { return ser::SaveMatrix(o, x); }},
                                            {"y", [this](std::ostream &o) { return ser::SaveMatrix(o, y); }},
                                            {"grad", [this](std::ostream &o) { return ser::SaveMatrix(o, grad); }},
                                            {"var_x", [this](std::ostream &o) { return ser::SaveMatrix(o, var_x); }},
                                            {"var_y", [this](std::ostream &o) { return ser::SaveMatrix(o, var_y); }},
                                            {"var_grad", [this](std::ostream &o) { return ser::SaveMatrix(o, var_grad); }},
                                            {"grad_flag", [this](std::ostream &o) { return ser::SaveMatrix(o, grad_flag); }}});
            }

            [[nodiscard]] bool
            Read(std::istream &s) {  // src/noisy_input_gp.cpp:519-606
                namespace ser = b200::serialization;
                return ser::ReadTokens(s, {{"x_dim", ser::ScalarReader(x_dim)},
                                           {"y_dim", ser::ScalarReader(y_dim)},
                                           {"num_samples", ser::ScalarReader(num_samples)},
                                           {"num_samples_with_grad", ser::ScalarReader(num_samples_with_grad)},
                                           {"x", [this](std::istream &i) { return ser::LoadMatrix(i, x); }},
                                           {"y", [this](std::istream &i) { return ser::LoadMatrix(i, y); }},
                                           {"grad", [this](std::istream &i) { return ser::LoadMatrix(i, grad); }},
                                           {"var_x", [this](std::istream &i) { return ser::LoadVector(i, var_x); }},
                                           {"var_y", [this](std::istream &i) { return ser::LoadVector(i, var_y); }},
                                           {"var_grad", [this](std::istream &i) { return ser::LoadVector(i, var_grad); }},
                                           {"grad_flag", [this](std::istream &i) { return ser::LoadVector(i, grad_flag); }}});
            }
        };

        class TestResult {  // noisy_input_gp.hpp:39-164
        protected:
            const NoisyInputGaussianProcess *m_gp_;
            const long m_num_test_;
            const bool m_support_gradient_;
            const long m_x_dim_;
            const long m_y_dim_;
            MatrixX m_x_test_;
            mutable MatrixX m_mean_;      // num_test x y_dim
            mutable MatrixX m_grad_;      // x_dim x (num_test * y_dim)
            mutable VectorX m_var_;       // num_test
            mutable MatrixX m_grad_var_;  // x_dim x num_test
            mutable MatrixX m_cov_;       // x_dim (x_dim + 1) / 2 x num_test

            void
            PrepareMean() const {
                if (m_mean_.size() > 0) { return; }
                m_mean_.resize(m_num_test_, m_y_dim_);
                if (m_support_gradient_) { m_grad_.resize(m_x_dim_, m_num_test_ * m_y_dim_); }
                m_gp_->m_ctx_->Check(Api::noisy_test(m_gp_->m_handle_, m_num_test_, m_x_test_.data(), m_x_test_.rows(), m_support_gradient_ ? 1 : 0, m_mean_.data(),
                                                     m_support_gradient_ ? m_grad_.data() : nullptr, nullptr, nullptr, nullptr),
                                     "erl_gp_noisy_test");
            }

            void
            PrepareAlphaTest() const {  // src/noisy_input_gp.cpp:362-376: everything that needs L^-1 Ktest, computed once
                if (m_var_.size() > 0) { return; }
                m_var_.resize(m_num_test_);
                if (m_support_gradient_) {
                    m_grad_var_.resize(m_x_dim_, m_num_test_);
                    m_cov_.resize(m_x_dim_ * (m_x_dim_ + 1) / 2, m_num_test_);
                }
                m_gp_->m_ctx_->Check(Api::noisy_test(m_gp_->m_handle_, m_num_test_, m_x_test_.data(), m_x_test_.rows(), m_support_gradient_ ? 1 : 0, nullptr, nullptr, m_var_.data(),
                                                     m_support_gradient_ ? m_grad_var_.data() : nullptr, m_support_gradient_ ? m_cov_.data() : nullptr),
                                     "erl_gp_noisy_test");
            }

        public:
            TestResult(const NoisyInputGaussianProcess *gp, const Eigen::Ref<const MatrixX> &mat_x_test, const bool will_predict_gradient)
                : m_gp_(gp),
                  m_num_test_(mat_x_test.cols()),
                  m_support_gradient_(will_predict_gradient),
                  m_x_dim_(gp->m_train_set_.x_dim),
                  m_y_dim_(gp->m_train_set_.y_dim),
                  m_x_test_(mat_x_test) {}

            virtual ~TestResult() = default;

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
                return m_y_dim_;
            }

            void
            GetMean(const long y_index, Eigen::Ref<VectorX> vec_f_out, const bool parallel) const {  // :125-145
                (void) parallel;
                PrepareMean();
                for (long i = 0; i < m_num_test_; ++i) { vec_f_out[i] = m_mean_(i, y_index); }
            }

            void
            GetMean(const long index, const long y_index, Dtype &f) const {  // :147-166
                PrepareMean();
                f = m_mean_(index, y_index);
            }

            [[nodiscard]] Eigen::VectorXb
            GetGradient(const long y_index, Eigen::Ref<MatrixX> mat_grad_out, const bool parallel) const {  // :168-207
                (void) parallel;
                b200::AssertM(m_support_gradient_, "m_support_gradient_ = false, it should be true to call GetGradient().");
                PrepareMean();
                Eigen::VectorXb valid_gradients(m_num_test_);
                for (long index = 0; index < m_num_test_; ++index) {
                    bool ok = true;
                    for (long j = 0; j < m_x_dim_; ++j) {
                        const Dtype g = m_grad_(j, index + y_index * m_num_test_);
                        mat_grad_out(j, index) = g;
                        ok = ok && std::isfinite(g);
                    }
                    valid_gradients[index] = ok;  // (the reference leaves the `true` entries uninitialised, :193)
                }
                return valid_gradients;
            }

            bool
            GetGradient(const long index, const long y_index, Dtype *grad) const {  // :209-231
                b200::AssertM(m_support_gradient_, "m_support_gradient_ = false, it should be true to call GetGradient().");
                PrepareMean();
                for (long j = 0; j < m_x_dim_; ++j) {
                    grad[j] = m_grad_(j, index + y_index * m_num_test_);
                    if (!std::isfinite(grad[j])) { return false; }
                }
                return true;
            }

            void
            GetMeanVariance(Eigen::Ref<VectorX> vec_var_out, const bool parallel) const {  // :233-246
                (void) parallel;
                PrepareAlphaTest();
                for (long i = 0; i < m_num_test_; ++i) { vec_var_out[i] = m_var_[i]; }
            }

            void
            GetMeanVariance(const long index, Dtype &var) const {  // :248-256
                PrepareAlphaTest();
                var = m_var_[index];
            }

            void
            GetGradientVariance(Eigen::Ref<MatrixX> mat_var_out, const bool parallel) const {  // :258-277
                (void) parallel;
                b200::AssertM(m_support_gradient_, "m_support_gradient_ = false, it should be true to call GetGradient().");
                PrepareAlphaTest();
                for (long i = 0; i < m_num_test_; ++i) {
                    for (long j = 0; j < m_x_dim_; ++j) { mat_var_out(j, i) = m_grad_var_(j, i); }
                }
            }

            void
            GetGradientVariance(const long index, Dtype *var) const {  // :279-298
                b200::AssertM(m_support_gradient_, "m_support_gradient_ = false, it should be true to call GetGradient().");
                PrepareAlphaTest();
                for (long j = 0; j < m_x_dim_; ++j) { var[j] = m_grad_var_(j, index); }
            }

            void
            GetCovariance(Eigen::Ref<MatrixX> mat_cov_out, const bool parallel) const {  // :300-333
                (void) parallel;
                b200::AssertM(m_support_gradient_, "m_support_gradient_ = false, it should be true to call GetGradient().");
                PrepareAlphaTest();
                for (long i = 0; i < m_num_test_; ++i) {
                    for (long j = 0; j < m_cov_.rows(); ++j) { mat_cov_out(j, i) = m_cov_(j, i); }
                }
            }

            void
            GetCovariance(const long index, Dtype *cov) const {  // :335-360
                b200::AssertM(m_support_gradient_, "m_support_gradient_ = false, it should be true to call GetGradient().");
                PrepareAlphaTest();
                for (long j = 0; j < m_cov_.rows(); ++j) { cov[j] = m_cov_(j, index); }
            }
        };

    protected:
        std::shared_ptr<Setting> m_setting_ = nullptr;
        std::shared_ptr<b200::DeviceContext> m_ctx_ = nullptr;
        typename Api::Noisy *m_handle_ = nullptr;
        bool m_trained_ = false;
        bool m_trained_once_ = false;
        bool m_k_train_updated_ = false;
        long m_k_train_rows_ = 0;
        long m_k_train_cols_ = 0;
        Dtype m_three_over_scale_square_ = 0.0f;
        int m_llt_info_ = 0;
        mutable bool m_host_copy_valid_ = false;
        mutable MatrixX m_mat_k_train_, m_mat_l_, m_mat_alpha_;
        TrainSet m_train_set_;

    public:
        explicit NoisyInputGaussianProcess(std::shared_ptr<Setting> setting, std::shared_ptr<b200::DeviceContext> ctx = nullptr)
            : m_setting_(std::move(setting)),
              m_ctx_(ctx ? std::move(ctx) : b200::DeviceContext::Default()) {
            b200::AssertM(m_setting_ != nullptr, "setting should not be nullptr.");
            b200::AssertM(m_setting_->kernel != nullptr, "setting->kernel should not be nullptr.");
            m_ctx_->Check(Api::noisy_create(m_ctx_->Get(), &m_handle_), "erl_gp_noisy_create");
        }

        NoisyInputGaussianProcess(const NoisyInputGaussianProcess &) = delete;
        NoisyInputGaussianProcess &
        operator=(const NoisyInputGaussianProcess &) = delete;

        virtual ~NoisyInputGaussianProcess() { Api::noisy_destroy(m_handle_); }

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

        void
        Reset(const long max_num_samples, const long x_dim, const long y_dim) {  // src/noisy_input_gp.cpp:700-724
            b200::AssertM(max_num_samples > 0, "max_num_samples should be > 0.");
            b200::AssertM(x_dim > 0, "x_dim should be > 0.");
            b200::AssertM(y_dim > 0, "y_dim should be > 0.");
            b200::AssertM(m_setting_->kernel->x_dim == -1 || m_setting_->kernel->x_dim == x_dim, "x_dim should be " + std::to_string(m_setting_->kernel->x_dim) + ".");
            b200::AssertM(m_setting_->max_num_samples < 0 || max_num_samples <= m_setting_->max_num_samples,
                          "max_num_samples should be <= " + std::to_string(m_setting_->max_num_samples) + ".");
            m_train_set_.Reset(max_num_samples, x_dim, y_dim, m_setting_->no_gradient_observation);
            m_trained_ = false;
            m_k_train_updated_ = false;
            m_k_train_rows_ = 0;
            m_k_train_cols_ = 0;
            m_three_over_scale_square_ = 3.0f / (m_setting_->kernel->scale * m_setting_->kernel->scale);
            m_host_copy_valid_ = false;
        }

        [[nodiscard]] TrainSet &
        GetTrainSet() {
            return m_train_set_;
        }

        [[nodiscard]] const TrainSet &
        GetTrainSet() const {
            return m_train_set_;
        }

        [[nodiscard]] MatrixX
        GetKtrainSized() const {  // :766-770
            if (m_k_train_rows_ <= 0 || m_k_train_cols_ <= 0) { return {}; }
            Materialise();
            return m_mat_k_train_;
        }

        [[nodiscard]] MatrixX
        GetAlphaSized() const {  // :784-788
            if (m_k_train_rows_ <= 0) { return {}; }
            Materialise();
            return m_mat_alpha_;
        }

        [[nodiscard]] const MatrixX &
        GetKtrain() const {  // (m x m, not the reference's max-size buffer)
            Materialise();
            return m_mat_k_train_;
        }

        [[nodiscard]] const MatrixX &
        GetAlpha() const {
            Materialise();
            return m_mat_alpha_;
        }

        [[nodiscard]] const MatrixX &
        GetCholeskyDecomposition() const {
            Materialise();
            return m_mat_l_;
        }

        [[nodiscard]] int
        GetLltInfo() const {  // an addition: the reference ignores llt().info() (:891)
            Materialise();
            return m_llt_info_;
        }

        bool
        UpdateKtrain() {  // :807-880 (fused with Train() on the device: the Gram matrix is built by erl_gp_noisy_train)
            if (m_k_train_updated_) { return true; }
            return m_train_set_.num_samples > 0;
        }

        [[nodiscard]] virtual bool
        Train() {  // :882-899
            if (m_trained_) { return false; }  // "The model has been trained. Please reset the model before training."
            m_trained_ = m_trained_once_;
            auto &ts = m_train_set_;
            if (!m_k_train_updated_) {
                if (ts.num_samples <= 0) { return false; }  // :811-814
                const bool no_grad = m_setting_->no_gradient_observation;
                long ng = 0;
                if (no_grad) {
                    for (long i = 0; i < ts.num_samples; ++i) { ts.grad_flag[i] = 0; }  // :817
                } else {
                    for (long i = 0; i < ts.num_samples; ++i) { ng += ts.grad_flag[i] != 0; }
                    b200::AssertM(ng == ts.num_samples_with_grad, "grad_flag.head(num_samples).count() != num_samples_with_grad");  // :822-826
                }
                const int kernel = covariance::KernelFromTypeName(m_setting_->kernel_type);
                m_ctx_->Check(Api::noisy_train(m_handle_, kernel, m_setting_->kernel->scale, ts.x_dim, ts.y_dim, ts.num_samples, ts.x.data(), ts.x.rows(), ts.y.data(), ts.y.rows(),
                                               no_grad ? nullptr : ts.grad.data(), no_grad ? 0 : ts.grad.rows(), ts.var_x.data(), ts.var_y.data(), no_grad ? nullptr : ts.var_grad.data(),
                                               ts.grad_flag.data(), no_grad ? 1 : 0),
                              "erl_gp_noisy_train");
                m_k_train_rows_ = m_k_train_cols_ = ts.num_samples + ts.x_dim * ng;
                m_k_train_updated_ = true;
                m_host_copy_valid_ = false;
            }
            m_trained_once_ = true;
            m_trained_ = true;
            return true;
        }

        [[nodiscard]] virtual std::shared_ptr<TestResult>
        Test(const Eigen::Ref<const MatrixX> &mat_x_test, const bool predict_gradient) const {  // :901-907
            if (!m_trained_) { return nullptr; }
            return std::make_shared<TestResult>(this, mat_x_test, predict_gradient);
        }

        // ---- operator== / Write / Read (src/noisy_input_gp.cpp:909-1160): same tokens in the same order; see serialization.hpp ----
        [[nodiscard]] bool
        operator==(const NoisyInputGaussianProcess &other) const {
            namespace ser = b200::serialization;
            if (!ser::SameGpSetting(*m_setting_, *other.m_setting_) || m_setting_->no_gradient_observation != other.m_setting_->no_gradient_observation) { return false; }
            if (m_trained_ != other.m_trained_ || m_trained_once_ != other.m_trained_once_ || m_k_train_updated_ != other.m_k_train_updated_) { return false; }
            if (m_k_train_rows_ != other.m_k_train_rows_ || m_k_train_cols_ != other.m_k_train_cols_) { return false; }
            if (!(m_three_over_scale_square_ == other.m_three_over_scale_square_)) { return false; }
            if (m_train_set_ != other.m_train_set_) { return false; }
            if (!m_k_train_updated_) { return true; }
            Materialise();
            other.Materialise();
            return ser::SameTopLeft(m_mat_k_train_, other.m_mat_k_train_, m_k_train_rows_, m_k_train_cols_) &&
                   ser::SameTopLeft(m_mat_l_, other.m_mat_l_, m_k_train_rows_, m_k_train_cols_) && m_mat_alpha_.cols() == other.m_mat_alpha_.cols() &&
                   ser::SameTopLeft(m_mat_alpha_, other.m_mat_alpha_, m_k_train_cols_, m_mat_alpha_.cols());
        }

        [[nodiscard]] bool
        operator!=(const NoisyInputGaussianProcess &other) const {
            return !(*this == other);
        }

        [[nodiscard]] virtual bool
        Write(std::ostream &s) const {
            namespace ser = b200::serialization;
            Materialise();
            return ser::WriteTokens(s, {{"setting", [this](std::ostream &o) { return ser::WriteGpSetting(o, *m_setting_) && (o << ' ' << m_setting_->no_gradient_observation).good(); }},
                                        {"trained", ser::ScalarWriter(m_trained_)},
                                        {"trained_once", ser::ScalarWriter(m_trained_once_)},
                                        {"k_train_updated", ser::ScalarWriter(m_k_train_updated_)},
                                        {"k_train_rows", ser::ScalarWriter(m_k_train_rows_)},
                                        {"k_train_cols", ser::ScalarWriter(m_k_train_cols_)},
                                        {"three_over_scale_square", [this](std::ostream &o) { o.write(reinterpret_cast<const char *>(&m_three_over_scale_square_), sizeof(Dtype)); return o.good(); }},
                                        {"kernel", [](std::ostream &o) { o << true << '\n'; return o.good(); }},
                                        {"mat_k_train", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_k_train_); }},
                                        {"mat_l", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_l_); }},
                                        {"mat_alpha", [this](std::ostream &o) { return ser::SaveMatrix(o, m_mat_alpha_); }},
                                        {"train_set", [this](std::ostream &o) { return m_train_set_.Write(o); }}});
        }

        [[nodiscard]] virtual bool
        Read(std::istream &s) {
            namespace ser = b200::serialization;
            bool trained = false, trained_once = false, updated = false, has_kernel = false;
            long rows = 0, cols = 0;
            MatrixX k, l, alpha;
            const bool ok = ser::ReadTokens(
                s,
                {{"setting", [this](std::istream &i) { return ser::ReadGpSetting(i, *m_setting_) && !(i >> m_setting_->no_gradient_observation).fail(); }},
                 {"trained", ser::ScalarReader(trained)},
                 {"trained_once", ser::ScalarReader(trained_once)},
                 {"k_train_updated", ser::ScalarReader(updated)},
                 {"k_train_rows", ser::ScalarReader(rows)},
                 {"k_train_cols", ser::ScalarReader(cols)},
                 {"three_over_scale_square", [this](std::istream &i) { i.read(reinterpret_cast<char *>(&m_three_over_scale_square_), sizeof(Dtype)); return i.good(); }},
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
            const long m = m_k_train_rows_;
            m_mat_k_train_.resize(m, m);
            m_mat_l_.resize(m, m);
            m_mat_alpha_.resize(m, m_train_set_.y_dim);
            long rows = 0;
            int info = 0;
            m_ctx_->Check(Api::noisy_get(m_handle_, &rows, &info, m_mat_k_train_.data(), m, m_mat_l_.data(), m, m_mat_alpha_.data(), m), "erl_gp_noisy_get");
            const_cast<NoisyInputGaussianProcess *>(this)->m_llt_info_ = info;
            m_host_copy_valid_ = true;
        }
    };

    using NoisyInputGaussianProcessD = NoisyInputGaussianProcess<double>;
    using NoisyInputGaussianProcessF = NoisyInputGaussianProcess<float>;
}  // namespace erl::gaussian_process
