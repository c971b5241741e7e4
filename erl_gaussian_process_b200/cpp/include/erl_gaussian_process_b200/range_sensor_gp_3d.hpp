// RangeSensorGaussianProcess3D<Dtype> — drop-in host class over the C ABI
// (include/erl_gaussian_process/range_sensor_gp_3d.hpp, src/range_sensor_gp_3d.cpp).
#pragma once

#include "mapping.hpp"
#include "sensor_frames.hpp"
#include "vanilla_gp.hpp"

#include <sstream>
#include <string>
#include <tuple>

namespace erl::gaussian_process {

    template<typename Dtype>
    class RangeSensorGaussianProcess3D {
    public:
        using Gp = VanillaGaussianProcess<Dtype>;
        using MappingDtype = Mapping<Dtype>;
        using RangeSensorFrame = geometry::LidarFrame3D<Dtype>;
        using Matrix3 = Eigen::Matrix3<Dtype>;
        using Matrix3X = Eigen::Matrix3X<Dtype>;
        using Vector3 = Eigen::Vector3<Dtype>;
        using MatrixX = Eigen::MatrixX<Dtype>;
        using VectorX = Eigen::VectorX<Dtype>;
        using Api = b200::Api<Dtype>;

        struct Setting {  // defaults: include/erl_gaussian_process/range_sensor_gp_3d.hpp:31-74
            long row_group_size = 24;
            long row_overlap_size = 6;
            long row_margin = 0;
            long col_group_size = 8;
            long col_overlap_size = 2;
            long col_margin = 0;
            long min_num_samples_per_group = 32;
            Dtype init_variance = 1e6f;
            Dtype sensor_range_var = 0.01f;
            Dtype max_valid_range_var = 0.1f;
            Dtype occ_test_temperature = 30.0f;
            std::string sensor_frame_type = "erl::geometry::LidarFrame3D";
            std::shared_ptr<typename RangeSensorFrame::Setting> sensor_frame = std::make_shared<typename RangeSensorFrame::Setting>();
            std::shared_ptr<typename Gp::Setting> gp = std::make_shared<typename Gp::Setting>();
            std::shared_ptr<typename MappingDtype::Setting> mapping = []() {
                auto s = std::make_shared<typename MappingDtype::Setting>();
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
            TestResult(const RangeSensorGaussianProcess3D *gp, const Eigen::Ref<const Matrix3X> &directions, const bool directions_are_local, const bool un_map) {
                const long n = directions.cols();
                MatrixX coords(2, n);
                Eigen::VectorXb coords_ok(n);
                const RangeSensorFrame *frame = gp->m_sensor_frame_.get();
                for (long i = 0; i < n; ++i) {  // DirWorldToFrame + ComputeFrameCoords, src/range_sensor_gp_3d.cpp:81-85
                    Dtype dir[3] = {directions(0, i), directions(1, i), directions(2, i)};
                    if (!directions_are_local) {
                        Dtype local[3];
                        frame->DirWorldToFrame(dir, local);
                        dir[0] = local[0], dir[1] = local[1], dir[2] = local[2];
                    }
                    Dtype dist;
                    coords_ok[i] = frame->ComputeFrameCoords(dir, dist, &coords(0, i));
                }
                m_mean_.resize(n);
                m_var_.resize(n);
                m_valid_.resize(n);
                gp->m_ctx_->Check(Api::range3d_test(gp->m_handle_, coords.data(), b200::MaskData(coords_ok), n, un_map, m_mean_.data(), m_var_.data(), b200::MaskData(m_valid_)), "erl_gp_range3d_test");
            }

            [[nodiscard]] long
            GetNumTest() const {
                return m_mean_.size();
            }

            [[nodiscard]] Eigen::VectorXb
            GetMean(Eigen::Ref<VectorX> vec_f_out, const bool parallel) const {
                (void) parallel;
                for (long i = 0; i < m_mean_.size(); ++i) {
                    if (m_valid_[i]) { vec_f_out[i] = m_mean_[i]; }
                }
                return m_valid_;
            }

            [[nodiscard]] Eigen::VectorXb
            GetVariance(Eigen::Ref<VectorX> vec_var_out, const bool parallel) const {
                (void) parallel;
                for (long i = 0; i < m_var_.size(); ++i) {
                    if (m_valid_[i]) { vec_var_out[i] = m_var_[i]; }
                }
                return m_valid_;
            }
        };

        // What the reference's GetGps()(r, c) (a VanillaGaussianProcess, include/erl_gaussian_process/range_sensor_gp_3d.hpp:136)
        // exposes to the code that walks the partition grid: trained flag, number of samples, Cholesky factor, alpha.  The state
        // lives on the device; a view is materialised on the host on first access after Train() (erl_gp_range3d_get_gp).
        class PartitionGp {
            const RangeSensorGaussianProcess3D *m_owner_;
            long m_row_, m_col_;
            mutable bool m_loaded_ = false, m_trained_ = false;
            mutable long m_n_ = 0;
            mutable MatrixX m_mat_l_;
            mutable VectorX m_alpha_;

            void
            Load() const {
                if (m_loaded_) { return; }
                m_trained_ = m_owner_->GetGp(m_row_, m_col_, m_n_, m_mat_l_, m_alpha_);
                m_loaded_ = true;
            }

        public:
            PartitionGp(const RangeSensorGaussianProcess3D *owner, const long row, const long col)
                : m_owner_(owner),
                  m_row_(row),
                  m_col_(col) {}

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

        // stands in for Eigen::MatrixX<std::shared_ptr<Gp>> (column-major grid of partition GPs)
        class GpGrid {
            long m_rows_ = 0, m_cols_ = 0;
            std::vector<std::shared_ptr<PartitionGp>> m_data_;

        public:
            void
            Resize(const RangeSensorGaussianProcess3D *owner, const long rows, const long cols) {
                m_rows_ = rows, m_cols_ = cols;
                m_data_.clear();
                for (long c = 0; c < cols; ++c) {
                    for (long r = 0; r < rows; ++r) { m_data_.push_back(std::make_shared<PartitionGp>(owner, r, c)); }
                }
            }

            [[nodiscard]] long
            rows() const {
                return m_rows_;
            }

            [[nodiscard]] long
            cols() const {
                return m_cols_;
            }

            [[nodiscard]] long
            size() const {
                return m_rows_ * m_cols_;
            }

            [[nodiscard]] const std::shared_ptr<PartitionGp> &
            operator()(const long r, const long c) const {
                return m_data_[static_cast<std::size_t>(r + c * m_rows_)];
            }

            [[nodiscard]] const std::shared_ptr<PartitionGp> *
            data() const {
                return m_data_.data();
            }
        };

    protected:
        std::shared_ptr<Setting> m_setting_ = nullptr;
        std::shared_ptr<b200::DeviceContext> m_ctx_ = nullptr;
        GpGrid m_gps_;
        typename Api::Range3d *m_handle_ = nullptr;
        bool m_trained_ = false;
        std::vector<std::tuple<long, long, Dtype, Dtype>> m_row_partitions_, m_col_partitions_;
        std::shared_ptr<RangeSensorFrame> m_sensor_frame_ = nullptr;
        std::shared_ptr<MappingDtype> m_mapping_ = nullptr;

    public:
        explicit RangeSensorGaussianProcess3D(std::shared_ptr<Setting> setting, std::shared_ptr<b200::DeviceContext> ctx = nullptr)
            : m_setting_(std::move(setting)),
              m_ctx_(ctx ? std::move(ctx) : b200::DeviceContext::Default()),
              m_sensor_frame_(std::make_shared<RangeSensorFrame>(m_setting_->sensor_frame)),
              m_mapping_(MappingDtype::Create(m_setting_->mapping)) {
            b200::AssertM(m_setting_->row_overlap_size % 2 == 0, "row_overlap_size must be even.");  // src/range_sensor_gp_3d.cpp:190-193
            b200::AssertM(m_setting_->col_overlap_size % 2 == 0, "col_overlap_size must be even.");  // :194-197
            m_setting_->gp->max_num_samples = m_setting_->row_group_size * m_setting_->col_group_size;  // :213
            m_setting_->gp->kernel->x_dim = 2;                                                        // :214
            erl_gp_range3d_setting s{};
            s.row_group_size = m_setting_->row_group_size, s.row_overlap_size = m_setting_->row_overlap_size, s.row_margin = m_setting_->row_margin;
            s.col_group_size = m_setting_->col_group_size, s.col_overlap_size = m_setting_->col_overlap_size, s.col_margin = m_setting_->col_margin;
            s.min_num_samples_per_group = m_setting_->min_num_samples_per_group;
            s.sensor_range_var = m_setting_->sensor_range_var;
            s.kernel = covariance::KernelFromTypeName(m_setting_->gp->kernel_type);
            s.kernel_scale = m_setting_->gp->kernel->scale;
            s.mapping = static_cast<int>(m_setting_->mapping->type);
            s.mapping_scale = m_setting_->mapping->scale;
            m_ctx_->Check(Api::range3d_create(m_ctx_->Get(), &s, m_sensor_frame_->GetFrameCoordsData(), m_sensor_frame_->Rows(), m_sensor_frame_->Cols(), &m_handle_), "erl_gp_range3d_create");
            long nr = 0, nc = 0;
            m_ctx_->Check(Api::range3d_grid(m_handle_, &nr, &nc), "erl_gp_range3d_grid");
            m_gps_.Resize(this, nr, nc);
            for (int axis = 0; axis < 2; ++axis) {
                const long num = axis == 0 ? nr : nc;
                std::vector<long> il(num), ir(num);
                std::vector<Dtype> cl(num), cr(num);
                m_ctx_->Check(Api::range3d_partitions(m_handle_, axis, il.data(), ir.data(), cl.data(), cr.data()), "erl_gp_range3d_partitions");
                auto &dst = axis == 0 ? m_row_partitions_ : m_col_partitions_;
                for (long i = 0; i < num; ++i) { dst.emplace_back(il[i], ir[i], cl[i], cr[i]); }
            }
        }

        RangeSensorGaussianProcess3D(const RangeSensorGaussianProcess3D &) = delete;
        RangeSensorGaussianProcess3D &
        operator=(const RangeSensorGaussianProcess3D &) = delete;

        ~RangeSensorGaussianProcess3D() { Api::range3d_destroy(m_handle_); }

        [[nodiscard]] bool
        IsTrained() const {
            return m_trained_;
        }

        [[nodiscard]] std::shared_ptr<const Setting>
        GetSetting() const {
            return m_setting_;
        }

        [[nodiscard]] const GpGrid &
        GetGps() const {  // include/erl_gaussian_process/range_sensor_gp_3d.hpp:136
            return m_gps_;
        }

        [[nodiscard]] const std::vector<std::tuple<long, long, Dtype, Dtype>> &
        GetRowPartitions() const {
            return m_row_partitions_;
        }

        [[nodiscard]] const std::vector<std::tuple<long, long, Dtype, Dtype>> &
        GetColPartitions() const {
            return m_col_partitions_;
        }

        [[nodiscard]] std::shared_ptr<const RangeSensorFrame>
        GetSensorFrame() const {
            return m_sensor_frame_;
        }

        void
        Reset() {
            m_trained_ = false;
            for (long i = 0; i < m_gps_.size(); ++i) { m_gps_.data()[i]->Invalidate(); }
        }

        [[nodiscard]] bool
        GetGp(const long row_part, const long col_part, long &n, MatrixX &mat_l, VectorX &alpha) const {
            const long max_n = m_setting_->row_group_size * m_setting_->col_group_size;
            int info = -1;
            mat_l.resize(max_n, max_n);
            alpha.resize(max_n);
            m_ctx_->Check(Api::range3d_get_gp(m_handle_, row_part, col_part, &info, &n, mat_l.data(), max_n, alpha.data()), "erl_gp_range3d_get_gp");
            return info == 0;
        }

        [[nodiscard]] bool
        Train(const Matrix3 &rotation, const Vector3 &translation, MatrixX ranges) {  // src/range_sensor_gp_3d.cpp:321-364
            Reset();
            m_sensor_frame_->UpdateRanges(rotation, translation, std::move(ranges));
            if (!m_sensor_frame_->IsValid()) { return false; }
            m_ctx_->Check(Api::range3d_train(m_handle_, m_sensor_frame_->GetRanges().data(), b200::MaskData(m_sensor_frame_->GetHitMask())), "erl_gp_range3d_train");
            m_trained_ = true;
            return true;
        }

        [[nodiscard]] std::shared_ptr<TestResult>
        Test(const Eigen::Ref<const Matrix3X> &directions, const bool directions_are_local, const bool un_map) const {
            if (!m_trained_) { return nullptr; }
            return std::make_shared<TestResult>(this, directions, directions_are_local, un_map);
        }

        [[nodiscard]] bool
        ComputeOcc(const Vector3 &pos_local, Dtype &dist_pos, Dtype &range_pred, Dtype &occ) const {  // src/range_sensor_gp_3d.cpp:409-439
            if (!m_trained_) { return false; }
            Dtype coords[2] = {0, 0};
            unsigned char coords_ok = m_sensor_frame_->ComputeFrameCoords(pos_local.data(), dist_pos, coords) ? 1 : 0;  // + CoordsIsInFrame: the partition search rejects the rest
            unsigned char ok = 0;
            m_ctx_->Check(Api::range3d_compute_occ(m_handle_, coords, &coords_ok, &dist_pos, 1, m_setting_->max_valid_range_var, m_setting_->occ_test_temperature, &range_pred, &occ, &ok),
                          "erl_gp_range3d_compute_occ");
            return ok != 0;
        }

        // ---- operator== / Write / Read (src/range_sensor_gp_3d.cpp:441-655): tokens setting, trained, gps, row_partitions, col_partitions,
        // sensor_frame, mapped_distances in the reference's order; framing in serialization.hpp.  Read() needs an object constructed
        // with the same Setting, restores the frame and replays Train() on the device; the stored tables and partition GPs are the check.
        [[nodiscard]] bool
        operator==(const RangeSensorGaussianProcess3D &other) const {
            namespace ser = b200::serialization;
            if (SettingText() != other.SettingText() || m_trained_ != other.m_trained_) { return false; }
            if (m_row_partitions_ != other.m_row_partitions_ || m_col_partitions_ != other.m_col_partitions_ || m_gps_.size() != other.m_gps_.size()) { return false; }
            if (!m_trained_) { return true; }
            for (long i = 0; i < m_gps_.size(); ++i) {
                if (!ser::SamePartitionGp(*m_gps_.data()[i], *other.m_gps_.data()[i])) { return false; }
            }
            return true;
        }

        [[nodiscard]] bool
        operator!=(const RangeSensorGaussianProcess3D &other) const {
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
                      o << m_gps_.rows() << ' ' << m_gps_.cols() << '\n';
                      for (long i = 0; i < m_gps_.size(); ++i) {
                          const char has_gp = 1;
                          o.write(&has_gp, 1);
                          if (!ser::WritePartitionGp(o, *m_gps_.data()[i])) { return false; }
                      }
                      return o.good();
                  }},
                 {"row_partitions", [this](std::ostream &o) { return ser::WritePartitions(o, m_row_partitions_); }},
                 {"col_partitions", [this](std::ostream &o) { return ser::WritePartitions(o, m_col_partitions_); }},
                 {"sensor_frame", [this](std::ostream &o) { return ser::SaveMatrix(o, m_sensor_frame_->GetRotationMatrix()) && ser::SaveMatrix(o, m_sensor_frame_->GetRanges()); }},
                 {"mapped_distances",
                  [this](std::ostream &o) {
                      MatrixX mapped(m_sensor_frame_->GetRanges().rows(), m_sensor_frame_->GetRanges().cols());
                      for (long i = 0; i < mapped.size(); ++i) { mapped.data()[i] = m_mapping_->map(m_sensor_frame_->GetRanges().data()[i]); }
                      return ser::SaveMatrix(o, mapped);
                  }}});
        }

        [[nodiscard]] bool
        Read(std::istream &s) {
            namespace ser = b200::serialization;
            bool trained = false;
            std::string setting_text;
            std::vector<std::tuple<long, long, Dtype, Dtype>> row_parts, col_parts;
            long gp_rows = 0, gp_cols = 0;
            MatrixX rotation, ranges, mapped;
            std::streampos gps_pos;
            const bool ok = ser::ReadTokens(
                s,
                {{"setting", [&setting_text](std::istream &i) { return static_cast<bool>(std::getline(i, setting_text)); }},
                 {"trained", ser::ScalarReader(trained)},
                 {"gps",
                  [&](std::istream &i) {
                      i >> gp_rows >> gp_cols;
                      ser::SkipLine(i);
                      gps_pos = i.tellg();
                      if (i.fail() || gp_rows < 0 || gp_cols < 0 || gp_rows * gp_cols > (1 << 24)) { return false; }
                      for (long g = 0; g < gp_rows * gp_cols; ++g) {
                          char has_gp = 0;
                          i.read(&has_gp, 1);
                          if (has_gp && !ser::ReadAndComparePartitionGp(i, PartitionGp(this, 0, 0), false)) { return false; }
                      }
                      return i.good();
                  }},
                 {"row_partitions", [&row_parts](std::istream &i) { return ser::ReadPartitions(i, row_parts); }},
                 {"col_partitions", [&col_parts](std::istream &i) { return ser::ReadPartitions(i, col_parts); }},
                 {"sensor_frame", [&](std::istream &i) { return ser::LoadMatrix(i, rotation) && ser::LoadMatrix(i, ranges); }},
                 {"mapped_distances", [&mapped](std::istream &i) { return ser::LoadMatrix(i, mapped); }}});
            if (!ok || setting_text != SettingText() || row_parts != m_row_partitions_ || col_parts != m_col_partitions_ || gp_rows != m_gps_.rows() || gp_cols != m_gps_.cols()) { return false; }
            Reset();
            if (!trained) { return true; }
            Eigen::VectorX<Dtype> translation(3);
            translation[0] = translation[1] = translation[2] = 0;
            if (!Train(rotation, translation, ranges)) { return false; }
            const std::streampos after = s.tellg();
            s.seekg(gps_pos);
            for (long g = 0; g < m_gps_.size(); ++g) {
                char has_gp = 0;
                s.read(&has_gp, 1);
                if (has_gp && !ser::ReadAndComparePartitionGp(s, *m_gps_.data()[g], true)) { return false; }
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
            o << t.row_group_size << ' ' << t.row_overlap_size << ' ' << t.row_margin << ' ' << t.col_group_size << ' ' << t.col_overlap_size << ' ' << t.col_margin << ' '
              << t.min_num_samples_per_group << ' ' << t.init_variance << ' ' << t.sensor_range_var << ' ' << t.max_valid_range_var << ' ' << t.occ_test_temperature << ' '
              << m_sensor_frame_->Rows() << ' ' << m_sensor_frame_->Cols() << ' ' << t.sensor_frame->valid_range_min << ' ' << t.sensor_frame->valid_range_max << ' '
              << static_cast<int>(t.mapping->type) << ' ' << t.mapping->scale << ' ' << t.gp->kernel->scale << ' ' << t.gp->kernel_type;
            return o.str();
        }

    public:
        // batched form for the per-voxel caller pattern: pos_local is 3 x T
        [[nodiscard]] Eigen::VectorXb
        ComputeOcc(const Matrix3X &pos_local, VectorX &dist_pos, VectorX &range_pred, VectorX &occ) const {
            const long n = pos_local.cols();
            Eigen::VectorXb ok(n);
            ok.setZero();
            if (!m_trained_) { return ok; }
            dist_pos.resize(n), range_pred.resize(n), occ.resize(n);
            MatrixX coords(2, n);
            Eigen::VectorXb coords_ok(n);
            for (long i = 0; i < n; ++i) { coords_ok[i] = m_sensor_frame_->ComputeFrameCoords(&pos_local(0, i), dist_pos[i], &coords(0, i)) ? 1 : 0; }
            m_ctx_->Check(Api::range3d_compute_occ(m_handle_, coords.data(), b200::MaskData(coords_ok), dist_pos.data(), n, m_setting_->max_valid_range_var, m_setting_->occ_test_temperature,
                                                   range_pred.data(), occ.data(), b200::MaskData(ok)),
                          "erl_gp_range3d_compute_occ");
            return ok;
        }
    };

    using RangeSensorGaussianProcess3Dd = RangeSensorGaussianProcess3D<double>;
    using RangeSensorGaussianProcess3Df = RangeSensorGaussianProcess3D<float>;
}  // namespace erl::gaussian_process
