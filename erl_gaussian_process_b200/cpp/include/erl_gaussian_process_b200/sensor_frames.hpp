// Minimal stand-ins for the erl_geometry v0.2.0 sensor frames (LidarFrame2D, LidarFrame3D): just the members the
// GP classes call.  erl_geometry is outside the hot path; a build against the real
// package passes its frames' arrays to the same C ABI instead.
#pragma once

#include "eigen_shim.hpp"

#include <cmath>
#include <limits>
#include <memory>
#include <vector>

namespace erl::geometry {

    template<typename Dtype>
    class LidarFrame2D {
    public:
        using VectorX = Eigen::VectorX<Dtype>;
        using MatrixX = Eigen::MatrixX<Dtype>;

        struct Setting {
            Dtype valid_range_min = 0.0;
            Dtype valid_range_max = std::numeric_limits<Dtype>::infinity();
            Dtype angle_min = -M_PI;
            Dtype angle_max = M_PI;
            long num_rays = 360;
            bool discontinuity_detection = false;
            Dtype discontinuity_factor = 10;
        };

    protected:
        std::shared_ptr<Setting> m_setting_;
        VectorX m_angles_, m_ranges_;
        Eigen::VectorXb m_mask_hit_, m_mask_continuous_;
        MatrixX m_rotation_;
        VectorX m_translation_;
        long m_num_hit_ = 0;

    public:
        explicit LidarFrame2D(std::shared_ptr<Setting> setting)
            : m_setting_(std::move(setting)),
              m_angles_(m_setting_->num_rays),
              m_rotation_(2, 2),
              m_translation_(2) {
            const long n = m_setting_->num_rays;
            for (long i = 0; i < n; ++i) { m_angles_[i] = n > 1 ? m_setting_->angle_min + (m_setting_->angle_max - m_setting_->angle_min) * Dtype(i) / Dtype(n - 1) : m_setting_->angle_min; }
            m_rotation_.setZero();
            m_rotation_(0, 0) = m_rotation_(1, 1) = 1;
            m_translation_.setZero();
        }

        // replace the uniformly spaced angles by measured ones (sensor logs carry their own angles)
        void
        SetAnglesInFrame(const VectorX &angles) {
            m_angles_ = angles;
        }

        void
        UpdateRanges(const MatrixX &rotation, const VectorX &translation, VectorX ranges) {
            m_rotation_ = rotation;
            m_translation_ = translation;
            m_ranges_ = std::move(ranges);
            const long n = m_ranges_.size();
            m_mask_hit_.resize(n);
            m_mask_continuous_.resize(n);
            m_num_hit_ = 0;
            for (long i = 0; i < n; ++i) {
                const Dtype r = m_ranges_[i];
                m_mask_hit_[i] = std::isfinite(r) && r >= m_setting_->valid_range_min && r <= m_setting_->valid_range_max;
                m_mask_continuous_[i] = 1;
                m_num_hit_ += m_mask_hit_[i];
            }
            if (m_setting_->discontinuity_detection) {
                Dtype rolling = 0;
                for (long i = 1; i < n; ++i) {
                    const Dtype diff = std::abs(m_ranges_[i] - m_ranges_[i - 1]);
                    if (i > 1 && rolling > 0 && diff > m_setting_->discontinuity_factor * rolling) { m_mask_continuous_[i - 1] = m_mask_continuous_[i] = 0; }
                    rolling = i > 1 ? Dtype(0.9) * rolling + Dtype(0.1) * diff : diff;
                }
            }
        }

        [[nodiscard]] const VectorX &
        GetAnglesInFrame() const {
            return m_angles_;
        }

        [[nodiscard]] const VectorX &
        GetRanges() const {
            return m_ranges_;
        }

        [[nodiscard]] const Eigen::VectorXb &
        GetHitMask() const {
            return m_mask_hit_;
        }

        [[nodiscard]] const Eigen::VectorXb &
        GetContinuityMask() const {
            return m_mask_continuous_;
        }

        [[nodiscard]] const MatrixX &
        GetRotationMatrix() const {
            return m_rotation_;
        }

        [[nodiscard]] const VectorX &
        GetTranslationVector() const {
            return m_translation_;
        }

        [[nodiscard]] bool
        IsValid() const {
            return m_num_hit_ > 0;
        }
    };

    // azimuth x elevation ray grid: frame_coords(r, c) = (azimuth_r, elevation_c)
    template<typename Dtype>
    class LidarFrame3D {
    public:
        using MatrixX = Eigen::MatrixX<Dtype>;

        struct Setting {
            Dtype valid_range_min = 0.0;
            Dtype valid_range_max = std::numeric_limits<Dtype>::infinity();
            Dtype azimuth_min = -M_PI, azimuth_max = M_PI;
            long num_azimuth_lines = 360;
            Dtype elevation_min = -M_PI / 2, elevation_max = M_PI / 2;
            long num_elevation_lines = 181;
        };

    protected:
        std::shared_ptr<Setting> m_setting_;
        std::vector<Dtype> m_frame_coords_;  // (r + c * rows) * 2 + k, the layout of Eigen::MatrixX<Vector2>
        MatrixX m_ranges_, m_rotation_;
        Eigen::MatrixXb m_mask_hit_;
        long m_num_hit_ = 0;

    public:
        explicit LidarFrame3D(std::shared_ptr<Setting> setting)
            : m_setting_(std::move(setting)),
              m_rotation_(3, 3) {
            const long rows = m_setting_->num_azimuth_lines, cols = m_setting_->num_elevation_lines;
            m_frame_coords_.resize(static_cast<std::size_t>(2 * rows * cols));
            for (long c = 0; c < cols; ++c) {
                const Dtype el = cols > 1 ? m_setting_->elevation_min + (m_setting_->elevation_max - m_setting_->elevation_min) * Dtype(c) / Dtype(cols - 1) : m_setting_->elevation_min;
                for (long r = 0; r < rows; ++r) {
                    const Dtype az = rows > 1 ? m_setting_->azimuth_min + (m_setting_->azimuth_max - m_setting_->azimuth_min) * Dtype(r) / Dtype(rows - 1) : m_setting_->azimuth_min;
                    m_frame_coords_[2 * (r + c * rows)] = az;
                    m_frame_coords_[2 * (r + c * rows) + 1] = el;
                }
            }
            m_rotation_.setZero();
            m_rotation_(0, 0) = m_rotation_(1, 1) = m_rotation_(2, 2) = 1;
        }

        [[nodiscard]] long
        Rows() const {
            return m_setting_->num_azimuth_lines;
        }

        [[nodiscard]] long
        Cols() const {
            return m_setting_->num_elevation_lines;
        }

        [[nodiscard]] const Dtype *
        GetFrameCoordsData() const {
            return m_frame_coords_.data();
        }

        void
        UpdateRanges(const MatrixX &rotation, const Eigen::VectorX<Dtype> &translation, MatrixX ranges) {
            (void) translation;
            m_rotation_ = rotation;
            m_ranges_ = std::move(ranges);
            m_mask_hit_.resize(m_ranges_.rows(), m_ranges_.cols());
            m_num_hit_ = 0;
            for (long i = 0; i < m_ranges_.size(); ++i) {
                const Dtype r = m_ranges_.data()[i];
                m_mask_hit_.data()[i] = std::isfinite(r) && r >= m_setting_->valid_range_min && r <= m_setting_->valid_range_max;
                m_num_hit_ += m_mask_hit_.data()[i];
            }
        }

        [[nodiscard]] const MatrixX &
        GetRanges() const {
            return m_ranges_;
        }

        [[nodiscard]] const Eigen::MatrixXb &
        GetHitMask() const {
            return m_mask_hit_;
        }

        [[nodiscard]] const MatrixX &
        GetRotationMatrix() const {
            return m_rotation_;
        }

        [[nodiscard]] bool
        IsValid() const {
            return m_num_hit_ > 0;
        }

        // world -> frame: R^T d
        void
        DirWorldToFrame(const Dtype *d, Dtype *out) const {
            for (int i = 0; i < 3; ++i) { out[i] = m_rotation_(0, i) * d[0] + m_rotation_(1, i) * d[1] + m_rotation_(2, i) * d[2]; }
        }

        [[nodiscard]] bool
        ComputeFrameCoords(const Dtype *dir_local, Dtype &dist, Dtype *frame_coords) const {
            dist = std::sqrt(dir_local[0] * dir_local[0] + dir_local[1] * dir_local[1] + dir_local[2] * dir_local[2]);
            if (!(dist > 0)) { return false; }
            frame_coords[0] = std::atan2(dir_local[1], dir_local[0]);
            frame_coords[1] = std::asin(dir_local[2] / dist);
            return true;
        }
    };

}  // namespace erl::geometry
