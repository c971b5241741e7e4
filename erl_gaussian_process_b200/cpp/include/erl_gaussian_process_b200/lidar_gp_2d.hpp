// LidarGaussianProcess2D<Dtype> — drop-in host class over the C ABI (include/erl_gaussian_process/lidar_gp_2d.hpp,
// src/lidar_gp_2d.cpp).  Same Setting fields and defaults, Train(rotation, translation, ranges),
// Test(angles, angles_are_local, un_map) -> TestResult::GetMean / GetVariance returning the VectorXb validity
// mask, GetAnglePartitions, ComputeOcc.  The partition GPs live in one device batch (erl_gp_lidar2d_*).
#pragma once

#include "mapping.hpp"
#include "sensor_frames.hpp"
#include "vanilla_gp.hpp"

#include <cmath>
#include <sstream>
#include <string>
#include <tuple>

namespace erl::gaussian_process {

    template<typename Dtype>
    class LidarGaussianProcess2D {
    public:
        using Gp = VanillaGaussianProcess<Dtype>;
        using GpSetting = typename Gp::Setting;
        using MappingDtype = Mapping<Dtype>;
        using MappingSetting = typename MappingDtype::Setting;
        using Matrix2 = Eigen::Matrix2<Dtype>;
        using Vector2 = Eigen::Vector2<Dtype>;
        using MatrixX = Eigen::MatrixX<Dtype>;
        using VectorX = Eigen::VectorX<Dtype>;
        using LidarFrame2D = geometry::LidarFrame2D<Dtype>;
        using LidarFrameSetting = typename LidarFrame2D::Setting;
        using Api = b200::Api<Dtype>;

        struct Setting {  // defaults: include/erl_gaussian_process/lidar_gp_2d.hpp:28-62
            bool partition_on_hit_rays = false;
            bool symmetric_partitions = true;
            long group_size = 26;
            long overlap_size = 6;
            long margin = 1;
            Dtype init_variance = 1e6f;
            Dtype sensor_range_var = 0.01f;
            Dtype discontinuity_var = 10.0f;
            Dtype max_valid_range_var = 0.1f;
            Dtype occ_test_temperature = 30.0f;
            std::shared_ptr<LidarFrameSetting> sensor_frame = std::make_shared<LidarFrameSetting>();
            std::shared_ptr<GpSetting> gp = std::make_shared<GpSetting>();
            std::shared_ptr<MappingSetting> mapping = []() {
                auto s = std::make_shared<MappingSetting>();
                s->type = MappingType::kInverseSqrt;
                s->scale = 1.0;
                return s;
            }();
        };

        class TestResult {
        protected:
            VectorX m_mean_, m_var_;
            Eigen::VectorXb m_valid_;

        public:
            TestResult(const LidarGaussianProcess2D *gp, const Eigen::Ref<const VectorX> &angles, const bool angles_are_local, const bool un_map) {
                const long n = angles.size();
                m_mean_.resize(n);
                m_var_.resize(n);
                m_valid_.resize(n);
                gp->m_ctx_->Check(Api::lidar2d_test(gp->m_handle_, angles.data(), n, angles_are_local, un_map, m_mean_.data(), m_var_.data(), b200::MaskData(m_valid_)), "erl_gp_lidar2d_test");
            }

            [[nodiscard]] long
            GetNumTest() const {
                return m_mean_.size();
            }

            // invalid rays are left unwritten, as src/lidar_gp_2d.cpp:112,120
            [[nodiscard]] Eigen::VectorXb
            GetMean(Eigen::Ref<VectorX> vec_f_out, const bool parallel) const {
                (void) parallel;
                for (long i = 0; i < m_mean_.size(); ++i) {
                    if (m_valid_[i]) { vec_f_out[i] = m_mean_[i]; }
                }
                return m_valid_;
            }

            [[nodiscard]] bool
            GetMean(const long index, Dtype &f) const {
                if (!m_valid_[index]) { return false; }
                f = m_mean_[index];
                return true;
            }

            [[nodiscard]] Eigen::VectorXb
            GetVariance(Eigen::Ref<VectorX> vec_var_out, const bool parallel) const {
                (void) parallel;
                for (long i = 0; i < m_var_.size(); ++i) {
                    if (m_valid_[i]) { vec_var_out[i] = m_var_[i]; }
                }
                return m_valid_;
            }

            [[nodiscard]] bool
            GetVariance(const long index, Dtype &var) const {
                if (!m_valid_[index]) { return false; }
                var = m_var_[index];
                return true;
            }
        };

        // What the reference's GetGps()[p] (a VanillaGaussianProcess, include/erl_gaussian_process/lidar_gp_2d.hpp:132) exposes to
        // the code that walks the partitions (erl_sdf_mapping): trained flag, number of samples, Cholesky factor, alpha.  The
        // state lives on the device; a view is materialised on the host on first access after Train() (erl_gp_lidar2d_get_gp).
        class PartitionGp {
            const LidarGaussianProcess2D *m_owner_;
            long m_index_;
            mutable bool m_loaded_ = false, m_trained_ = false;
            mutable long m_n_ = 0;
            mutable MatrixX m_mat_l_;
            mutable VectorX m_alpha_;

            void
            Load() const {
                if (m_loaded_) { return; }
                m_trained_ = m_owner_->GetGp(m_index_, m_n_, m_mat_l_, m_alpha_);
                m_loaded_ = true;
            }

        public:
            PartitionGp(const LidarGaussianProcess2D *owner, const long index)
                : m_owner_(owner),
                  m_index_(index) {}

            void
            Invalidate() const {
                m_loaded_ = false;
            }

            [[nodiscard]] bool
            IsTrained() const {
                Load();
                return m_trained_;
            }

            [[nodiscard]] long
            GetNumTrainSamples() const {  // TrainSet::num_samples
                Load();
                return m_trained_ ? m_n_ : 0;
            }

            [[nodiscard]] const MatrixX &
            GetCholeskyDecomposition() const {  // max_num_samples x max_num_samples buffer, the factor in its leading n x n block
                Load();
                return m_mat_l_;
            }

            [[nodiscard]] const VectorX &
            GetAlpha() const {
                Load();
                return m_alpha_;
            }
        };

    protected:
        std::shared_ptr<Setting> m_setting_ = nullptr;
        std::shared_ptr<b200::DeviceContext> m_ctx_ = nullptr;
        std::vector<std::shared_ptr<PartitionGp>> m_gps_;
        typename Api::Lidar2d *m_handle_ = nullptr;
        bool m_trained_ = false;
        std::vector<std::tuple<long, long, Dtype, Dtype>> m_angle_partitions_;
        std::shared_ptr<LidarFrame2D> m_sensor_frame_ = nullptr;
        std::shared_ptr<MappingDtype> m_mapping_ = nullptr;

    public:
        explicit LidarGaussianProcess2D(std::shared_ptr<Setting> setting, std::shared_ptr<LidarFrame2D> sensor_frame = nullptr, std::shared_ptr<b200::DeviceContext> ctx = nullptr)
            : m_setting_(std::move(setting)),
              m_ctx_(ctx ? std::move(ctx) : b200::DeviceContext::Default()),
              m_sensor_frame_(sensor_frame ? std::move(sensor_frame) : std::make_shared<LidarFrame2D>(m_setting_->sensor_frame)),
              m_mapping_(MappingDtype::Create(m_setting_->mapping)) {
            const VectorX &angles = m_sensor_frame_->GetAnglesInFrame();
            if (angles.size() <= m_setting_->overlap_size) { return; }  // "no enough samples to perform partition", :177-180
            m_setting_->gp->max_num_samples = m_setting_->group_size;  // :249
            m_setting_->gp->kernel->x_dim = 1;                         // :250
            erl_gp_lidar2d_setting s{};
            s.symmetric_partitions = m_setting_->symmetric_partitions;
            s.group_size = m_setting_->group_size;
            s.overlap_size = m_setting_->overlap_size;
            s.margin = m_setting_->margin;
            s.sensor_range_var = m_setting_->sensor_range_var;
            s.discontinuity_var = m_setting_->discontinuity_var;
            s.discontinuity_detection = m_setting_->sensor_frame->discontinuity_detection;
            s.kernel = covariance::KernelFromTypeName(m_setting_->gp->kernel_type);
            s.kernel_scale = m_setting_->gp->kernel->scale;
            s.mapping = static_cast<int>(m_setting_->mapping->type);
            s.mapping_scale = m_setting_->mapping->scale;
            // PartitionOnHitRays (src/lidar_gp_2d.cpp:302-348): the table is empty until the first Train() and follows the hit rays of
            // every frame; the reference's out-of-range indices (:329-347) are clamped by the library
            s.partition_on_hit_rays = m_setting_->partition_on_hit_rays;
            m_ctx_->Check(Api::lidar2d_create(m_ctx_->Get(), &s, angles.data(), angles.size(), &m_handle_), "erl_gp_lidar2d_create");
            FetchPartitions();
        }

    protected:
        void
        FetchPartitions() {  // m_angle_partitions_ / m_gps_ as the library holds them (m_gps_.resize(num_groups), :317)
            long num = 0;
            m_ctx_->Check(Api::lidar2d_num_partitions(m_handle_, &num), "erl_gp_lidar2d_num_partitions");
            std::vector<long> il(num), ir(num);
            std::vector<Dtype> cl(num), cr(num);
            if (num > 0) { m_ctx_->Check(Api::lidar2d_partitions(m_handle_, il.data(), ir.data(), cl.data(), cr.data()), "erl_gp_lidar2d_partitions"); }
            m_angle_partitions_.clear();
            for (long i = 0; i < num; ++i) { m_angle_partitions_.emplace_back(il[i], ir[i], cl[i], cr[i]); }
            while (static_cast<long>(m_gps_.size()) > num) { m_gps_.pop_back(); }
            while (static_cast<long>(m_gps_.size()) < num) { m_gps_.push_back(std::make_shared<PartitionGp>(this, static_cast<long>(m_gps_.size()))); }
        }

    public:

        LidarGaussianProcess2D(const LidarGaussianProcess2D &) = delete;
        LidarGaussianProcess2D &
        operator=(const LidarGaussianProcess2D &) = delete;

        ~LidarGaussianProcess2D() { Api::lidar2d_destroy(m_handle_); }

        [[nodiscard]] bool
        IsTrained() const {
            return m_trained_;
        }

        [[nodiscard]] std::shared_ptr<Setting>
        GetSetting() const {
            return m_setting_;
        }

        [[nodiscard]] const std::vector<std::shared_ptr<PartitionGp>> &
        GetGps() const {  // include/erl_gaussian_process/lidar_gp_2d.hpp:132
            return m_gps_;
        }

        [[nodiscard]] const std::vector<std::tuple<long, long, Dtype, Dtype>> &
        GetAnglePartitions() const {
            return m_angle_partitions_;
        }

        [[nodiscard]] std::shared_ptr<const LidarFrame2D>
        GetSensorFrame() const {
            return m_sensor_frame_;
        }

        [[nodiscard]] std::shared_ptr<const MappingDtype>
        GetMapping() const {
            return m_mapping_;
        }

        void
        Reset() {
            m_trained_ = false;
            for (const auto &gp: m_gps_) { gp->Invalidate(); }
        }

        // partition GP p as the reference's GetGps()[p] exposes it: trained flag, n, L (n x n), alpha
        [[nodiscard]] bool
        GetGp(const long p, long &n, MatrixX &mat_l, VectorX &alpha) const {
            int info = -1;
            mat_l.resize(m_setting_->group_size, m_setting_->group_size);
            alpha.resize(m_setting_->group_size);
            m_ctx_->Check(Api::lidar2d_get_gp(m_handle_, p, &info, &n, mat_l.data(), m_setting_->group_size, alpha.data()), "erl_gp_lidar2d_get_gp");
            return info == 0;
        }

        [[nodiscard]] bool
        Train(const Matrix2 &rotation, const Vector2 &translation, VectorX ranges) {  // src/lidar_gp_2d.cpp:350-396
            Reset();
            if (m_handle_ == nullptr) { return false; }
            m_sensor_frame_->UpdateRanges(rotation, translation, std::move(ranges));
            if (!m_sensor_frame_->IsValid()) { return false; }
            m_ctx_->Check(
                Api::lidar2d_train(m_handle_, rotation.data(), m_sensor_frame_->GetRanges().data(), b200::MaskData(m_sensor_frame_->GetHitMask()), b200::MaskData(m_sensor_frame_->GetContinuityMask())),
                "erl_gp_lidar2d_train");
            if (m_setting_->partition_on_hit_rays) { FetchPartitions(); }  // :364
            m_trained_ = true;
            return true;
        }

        [[nodiscard]] std::shared_ptr<TestResult>
        Test(const Eigen::Ref<const VectorX> &angles, const bool angles_are_local, const bool un_map) const {
            if (!m_trained_) { return nullptr; }
            return std::make_shared<TestResult>(this, angles, angles_are_local, un_map);
        }

        [[nodiscard]] bool
        ComputeOcc(const Vector2 &pos_local, Dtype &dist_pos, Dtype &range_pred, Dtype &occ) const {  // src/lidar_gp_2d.cpp:428-459
            if (!m_trained_) { return false; }
            unsigned char ok = 0;
            m_ctx_->Check(
                Api::lidar2d_compute_occ(m_handle_, pos_local.data(), 1, m_setting_->max_valid_range_var, m_setting_->occ_test_temperature, &dist_pos, &range_pred, &occ, &ok),
                "erl_gp_lidar2d_compute_occ");
            return ok != 0;
        }

        // ---- operator== / Write / Read (src/lidar_gp_2d.cpp:461-635): tokens setting, trained, gps, angle_partitions, sensor_frame,
        // mapped_distances in the reference's order; framing in serialization.hpp.  Read() needs an object constructed with the same
        // Setting (the stream's copy is checked against it), restores the frame and replays Train() on the device; the stored partition
        // table and partition GPs are the check.
        [[nodiscard]] bool
        operator==(const LidarGaussianProcess2D &other) const {
            namespace ser = b200::serialization;
            if (SettingText() != other.SettingText() || m_trained_ != other.m_trained_) { return false; }
            if (m_angle_partitions_ != other.m_angle_partitions_ || m_gps_.size() != other.m_gps_.size()) { return false; }
            if (!m_trained_) { return true; }
            for (std::size_t i = 0; i < m_gps_.size(); ++i) {
                if (!ser::SamePartitionGp(*m_gps_[i], *other.m_gps_[i])) { return false; }
            }
            const auto &r0 = m_sensor_frame_->GetRanges(), &r1 = other.m_sensor_frame_->GetRanges();
            if (r0.size() != r1.size()) { return false; }
            for (long i = 0; i < r0.size(); ++i) {
                if (!(r0[i] == r1[i]) && !(std::isnan(r0[i]) && std::isnan(r1[i]))) { return false; }
            }
            return true;
        }

        [[nodiscard]] bool
        operator!=(const LidarGaussianProcess2D &other) const {
            return !(*this == other);
        }

        [[nodiscard]] bool
        Write(std::ostream &s) const {
            namespace ser = b200::serialization;
            return ser::WriteTokens(
                s,
                {{"setting", [this](std::ostream &o) { o << SettingText() << '\n'; return o.good(); }},  // (a whole line: the kernel type name may hold blanks)
                 {"trained", ser::ScalarWriter(m_trained_)},
                 {"gps",
                  [this](std::ostream &o) {
                      o << m_gps_.size() << '\n';
                      for (const auto &g: m_gps_) {
                          const char has_gp = 1;
                          o.write(&has_gp, 1);
                          if (!ser::WritePartitionGp(o, *g)) { return false; }
                      }
                      return o.good();
                  }},
                 {"angle_partitions", [this](std::ostream &o) { return ser::WritePartitions(o, m_angle_partitions_); }},
                 {"sensor_frame",
                  [this](std::ostream &o) {
                      return ser::SaveMatrix(o, m_sensor_frame_->GetRotationMatrix()) && ser::SaveMatrix(o, m_sensor_frame_->GetTranslationVector()) &&
                             ser::SaveMatrix(o, m_sensor_frame_->GetRanges());
                  }},
                 {"mapped_distances",
                  [this](std::ostream &o) {
                      VectorX mapped(m_sensor_frame_->GetRanges().size());
                      for (long i = 0; i < mapped.size(); ++i) { mapped[i] = m_mapping_->map(m_sensor_frame_->GetRanges()[i]); }  // m_mapped_distances_, :234
                      return ser::SaveMatrix(o, mapped);
                  }}});
        }

        [[nodiscard]] bool
        Read(std::istream &s) {
            namespace ser = b200::serialization;
            bool trained = false;
            std::string setting_text;
            std::vector<std::tuple<long, long, Dtype, Dtype>> parts;
            std::string gps_bytes;
            std::size_t num_gps = 0;
            MatrixX rotation;
            VectorX translation, ranges, mapped;
            std::streampos gps_pos;
            const bool ok = ser::ReadTokens(
                s,
                {{"setting", [&setting_text](std::istream &i) { return static_cast<bool>(std::getline(i, setting_text)); }},
                 {"trained", ser::ScalarReader(trained)},
                 {"gps",
                  [&](std::istream &i) {  // parsed now (to find the end of the token), compared after the training has been replayed
                      i >> num_gps;
                      ser::SkipLine(i);
                      gps_pos = i.tellg();
                      if (i.fail() || num_gps > (1u << 24)) { return false; }
                      for (std::size_t g = 0; g < num_gps; ++g) {
                          char has_gp = 0;
                          i.read(&has_gp, 1);
                          if (has_gp && !ser::ReadAndComparePartitionGp(i, PartitionGp(this, 0), false)) { return false; }
                      }
                      return i.good();
                  }},
                 {"angle_partitions", [&parts](std::istream &i) { return ser::ReadPartitions(i, parts); }},
                 {"sensor_frame", [&](std::istream &i) { return ser::LoadMatrix(i, rotation) && ser::LoadVector(i, translation) && ser::LoadVector(i, ranges); }},
                 {"mapped_distances", [&mapped](std::istream &i) { return ser::LoadVector(i, mapped); }}});
            if (!ok || setting_text != SettingText()) { return false; }
            Reset();
            if (!trained) { return m_setting_->partition_on_hit_rays || parts == m_angle_partitions_; }
            if (!Train(rotation, translation, ranges) || parts != m_angle_partitions_ || num_gps != m_gps_.size()) { return false; }
            const std::streampos after = s.tellg();
            s.seekg(gps_pos);
            for (std::size_t g = 0; g < num_gps; ++g) {
                char has_gp = 0;
                s.read(&has_gp, 1);
                if (has_gp && !ser::ReadAndComparePartitionGp(s, *m_gps_[g], true)) { return false; }
            }
            s.seekg(after);
            return s.good();
        }

    protected:
        [[nodiscard]] std::string
        SettingText() const {  // one line: every Setting field that shapes the model (the reference writes its Setting as YAML)
            std::ostringstream o;
            o.precision(17);
            const auto &t = *m_setting_;
            o << t.partition_on_hit_rays << ' ' << t.symmetric_partitions << ' ' << t.group_size << ' ' << t.overlap_size << ' ' << t.margin << ' ' << t.init_variance << ' ' << t.sensor_range_var << ' '
              << t.discontinuity_var << ' ' << t.max_valid_range_var << ' ' << t.occ_test_temperature << ' ' << t.sensor_frame->num_rays << ' ' << t.sensor_frame->angle_min << ' '
              << t.sensor_frame->angle_max << ' ' << t.sensor_frame->valid_range_min << ' ' << t.sensor_frame->valid_range_max << ' ' << t.sensor_frame->discontinuity_detection << ' '
              << static_cast<int>(t.mapping->type) << ' ' << t.mapping->scale << ' ' << t.gp->kernel->scale << ' ' << t.gp->kernel_type;
            return o.str();
        }

    public:
        // batched form of ComputeOcc for the per-voxel caller pattern: pos_local is 2 x T
        [[nodiscard]] Eigen::VectorXb
        ComputeOcc(const MatrixX &pos_local, VectorX &dist_pos, VectorX &range_pred, VectorX &occ) const {
            const long n = pos_local.cols();
            Eigen::VectorXb ok(n);
            ok.setZero();
            if (!m_trained_) { return ok; }
            dist_pos.resize(n), range_pred.resize(n), occ.resize(n);
            m_ctx_->Check(Api::lidar2d_compute_occ(m_handle_, pos_local.data(), n, m_setting_->max_valid_range_var, m_setting_->occ_test_temperature, dist_pos.data(), range_pred.data(),
                                                   occ.data(), b200::MaskData(ok)),
                          "erl_gp_lidar2d_compute_occ");
            return ok;
        }
    };

    using LidarGaussianProcess2Dd = LidarGaussianProcess2D<double>;
    using LidarGaussianProcess2Df = LidarGaussianProcess2D<float>;
}  // namespace erl::gaussian_process
