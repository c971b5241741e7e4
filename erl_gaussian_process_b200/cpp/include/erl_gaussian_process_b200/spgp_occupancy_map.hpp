// SpGpOccupancyMap<Dtype, Dim> — drop-in host class over SparsePseudoInputGaussianProcess (include/erl_gaussian_process/
// spgp_occupancy_map.hpp + src/spgp_occupancy_map.cpp): an occupancy log-odds field regressed by an SPGP from a labelled
// point set (occupied = the hit points of a range scan, free = points sampled along the rays).
//
//   Update  (:82-124)   labels -> logodd_occupied / logodd_free, constant logodd_variance, SPGP Reset + Update
//   Predict (:126-152)  SPGP mean (+ gradient of the mean) at query points
//
// The dataset itself is produced by erl_geometry's OccupancyMap::GenerateDataset (:54-78), which is not part of the reference
// repository (an absent dependency).  `UpdateWithDataset` takes a dataset produced by the caller's generator; `GenerateDataset`
// below is a STAND-IN with the documented semantics of the Setting fields (min / max distance, free_points_per_meter,
// free_sampling_margin, map boundary, max_dataset_size), not a restatement of erl_geometry's sampler: its random stream and
// sample order are its own.
#pragma once

#include "sparse_pseudo_input_gp.hpp"

#include <cmath>
#include <cstdint>
#include <istream>
#include <memory>
#include <ostream>
#include <random>
#include <utility>
#include <vector>

namespace erl::gaussian_process {

    // axis-aligned box, the two fields of erl_geometry's Aabb that this class reads and serializes (:194-204)
    template<typename Dtype>
    struct Aabb {
        Eigen::VectorX<Dtype> center;
        Eigen::VectorX<Dtype> half_sizes;

        [[nodiscard]] bool
        contains(const Dtype *p, const long dim) const {
            for (long d = 0; d < dim; ++d) {
                if (std::abs(p[d] - center[d]) > half_sizes[d]) { return false; }
            }
            return true;
        }
    };

    template<typename Dtype, int Dim>
    class SpGpOccupancyMap {
    public:
        using SpGp = SparsePseudoInputGaussianProcess<Dtype>;
        using SpGpSetting = typename SpGp::Setting;
        using MatrixX = Eigen::MatrixX<Dtype>;
        using VectorX = Eigen::VectorX<Dtype>;
        using MatrixDX = Eigen::MatrixX<Dtype>;  // Dim x n
        using VectorD = Eigen::VectorX<Dtype>;   // Dim
        using AabbD = Aabb<Dtype>;

        struct Setting {  // spgp_occupancy_map.hpp:20-38
            std::shared_ptr<SpGpSetting> sp_gp = std::make_shared<SpGpSetting>();
            Dtype min_distance = 0.5f;
            Dtype max_distance = 30.0f;
            Dtype free_points_per_meter = 2.0f;
            Dtype free_sampling_margin = 0.05f;
            bool parallel = true;
            Dtype logodd_free = -5.0f;
            Dtype logodd_occupied = 5.0f;
            Dtype logodd_variance = 0.0001f;
        };

        SpGpOccupancyMap() = delete;

        SpGpOccupancyMap(std::shared_ptr<Setting> setting, MatrixX pseudo_points, AabbD map_boundary, const uint64_t seed)  // :43-52
            : m_setting_(CheckedSetting(std::move(setting))),
              m_sp_gp_(m_setting_->sp_gp, std::move(pseudo_points)),
              m_map_boundary_(std::move(map_boundary)),
              m_generator_(seed) {
            b200::AssertM(m_sp_gp_.GetPseudoPoints().rows() == Dim, "pseudo_points should have Dim rows");
        }

        [[nodiscard]] const SpGp &
        GetSpGp() const {
            return m_sp_gp_;
        }

        [[nodiscard]] std::shared_ptr<const Setting>
        GetSetting() const {
            return m_setting_;
        }

        // Stand-in for geometry::OccupancyMap::GenerateDataset (see the header comment): per selected point at distance r from the
        // sensor with min_distance <= r <= max_distance, the point itself (label 1) if inside the map boundary and
        // floor(r * free_points_per_meter) free points (label 0) at uniformly drawn ray fractions in [margin, 1 - margin];
        // stops when max_dataset_size samples exist (<= 0: no limit).
        void
        GenerateDataset(
            const Eigen::Ref<const VectorD> &sensor_position,
            const Eigen::Ref<const MatrixDX> &points,
            const std::vector<long> &point_indices,
            const long max_dataset_size,
            long &num_samples,
            MatrixDX &dataset_points,
            VectorX &dataset_labels,
            std::vector<long> &hit_indices) {
            const long num_points = point_indices.empty() ? points.cols() : static_cast<long>(point_indices.size());
            std::vector<Dtype> xs;
            std::vector<Dtype> labels;
            hit_indices.clear();
            std::uniform_real_distribution<Dtype> fraction(m_setting_->free_sampling_margin, Dtype(1) - m_setting_->free_sampling_margin);
            auto full = [&]() { return max_dataset_size > 0 && static_cast<long>(labels.size()) >= max_dataset_size; };
            for (long k = 0; k < num_points && !full(); ++k) {
                const long idx = point_indices.empty() ? k : point_indices[static_cast<std::size_t>(k)];
                Dtype v[Dim], p[Dim];
                Dtype r2 = 0;
                bool finite = true;
                for (int d = 0; d < Dim; ++d) {
                    p[d] = points(d, idx);
                    v[d] = p[d] - sensor_position[d];
                    r2 += v[d] * v[d];
                    finite = finite && std::isfinite(p[d]);
                }
                const Dtype r = std::sqrt(r2);
                if (!finite || r < m_setting_->min_distance || r > m_setting_->max_distance) { continue; }
                if (m_map_boundary_.contains(p, Dim)) {
                    xs.insert(xs.end(), p, p + Dim);
                    labels.push_back(1);
                    hit_indices.push_back(idx);
                }
                const long num_free = static_cast<long>(std::floor(r * m_setting_->free_points_per_meter));
                for (long f = 0; f < num_free && !full(); ++f) {
                    const Dtype t = fraction(m_generator_);
                    Dtype q[Dim];
                    for (int d = 0; d < Dim; ++d) { q[d] = sensor_position[d] + t * v[d]; }
                    if (!m_map_boundary_.contains(q, Dim)) { continue; }
                    xs.insert(xs.end(), q, q + Dim);
                    labels.push_back(0);
                }
            }
            num_samples = static_cast<long>(labels.size());
            if (dataset_points.rows() != Dim || dataset_points.cols() < num_samples) { dataset_points.resize(Dim, num_samples); }
            if (dataset_labels.size() < num_samples) { dataset_labels.resize(num_samples); }
            for (long i = 0; i < num_samples; ++i) {
                for (int d = 0; d < Dim; ++d) { dataset_points(d, i) = xs[static_cast<std::size_t>(i * Dim + d)]; }
                dataset_labels[i] = labels[static_cast<std::size_t>(i)];
            }
        }

        bool
        Update(
            const Eigen::Ref<const VectorD> &sensor_position,
            const Eigen::Ref<const MatrixDX> &points,
            const std::vector<long> &point_indices,
            long &num_samples,
            MatrixDX &dataset_points,
            VectorX &dataset_labels,
            std::vector<long> &hit_indices) {  // :82-124
            const long max_dataset_size = m_setting_->sp_gp->max_num_samples;
            b200::AssertM(max_dataset_size > 0, "max_dataset_size should be greater than 0");
            GenerateDataset(sensor_position, points, point_indices, max_dataset_size, num_samples, dataset_points, dataset_labels, hit_indices);
            return UpdateWithDataset(num_samples, dataset_points, dataset_labels);
        }

        // the part of Update() after the dataset exists (:106-123): for callers that run erl_geometry's generator themselves
        bool
        UpdateWithDataset(const long num_samples, const MatrixDX &dataset_points, const VectorX &dataset_labels) {
            if (num_samples == 0) { return false; }  // "No valid points generated for update. Skipping update."
            m_sp_gp_.Reset(num_samples, Dim, 1);
            auto &train_set = m_sp_gp_.GetTrainSet();
            train_set.x_dim = Dim;
            train_set.y_dim = 1;
            train_set.num_samples = num_samples;
            for (long i = 0; i < num_samples; ++i) {
                for (int d = 0; d < Dim; ++d) { train_set.x(d, i) = dataset_points(d, i); }
                train_set.y(i, 0) = dataset_labels[i] > 0 ? m_setting_->logodd_occupied : m_setting_->logodd_free;
                train_set.var[i] = m_setting_->logodd_variance;
            }
            return m_sp_gp_.Update(m_setting_->parallel);
        }

        void
        Predict(const Eigen::Ref<const MatrixDX> &points, const bool compute_gradient, const bool parallel, VectorX &logodd, MatrixDX &gradient) const {  // :126-140
            auto test_result = m_sp_gp_.Test(points, compute_gradient);
            b200::AssertM(test_result != nullptr, "Predict() before the first successful Update()");
            if (logodd.size() < points.cols()) { logodd.resize(points.cols()); }
            test_result->GetMean(0, logodd, parallel);
            if (compute_gradient) {
                if (gradient.rows() != Dim || gradient.cols() < points.cols()) { gradient.resize(Dim, points.cols()); }
                (void) test_result->GetGradient(0, gradient, parallel);
            }
        }

        void
        Predict(const VectorD &point, const bool compute_gradient, Dtype &logodd, VectorD &gradient) const {  // :142-152
            MatrixDX x;
            x.resize(Dim, 1);
            for (int d = 0; d < Dim; ++d) { x(d, 0) = point[d]; }
            auto test_result = m_sp_gp_.Test(x, compute_gradient);
            b200::AssertM(test_result != nullptr, "Predict() before the first successful Update()");
            test_result->GetMean(0, 0, logodd);
            if (compute_gradient) {
                if (gradient.size() < Dim) { gradient.resize(Dim); }
                (void) test_result->GetGradient(0, 0, gradient.data());
            }
        }

        void
        PredictGradient(const Eigen::Ref<const MatrixDX> &points, const bool parallel, MatrixDX &gradient) const {  // :154-162
            auto test_result = m_sp_gp_.Test(points, true);
            b200::AssertM(test_result != nullptr, "PredictGradient() before the first successful Update()");
            if (gradient.rows() != Dim || gradient.cols() < points.cols()) { gradient.resize(Dim, points.cols()); }
            (void) test_result->GetGradient(0, gradient, parallel);
        }

        [[nodiscard]] bool
        operator==(const SpGpOccupancyMap &other) const {  // :244-254: setting, SPGP, boundary (not the generator)
            const Setting &a = *m_setting_, &b = *other.m_setting_;
            if (a.min_distance != b.min_distance || a.max_distance != b.max_distance || a.free_points_per_meter != b.free_points_per_meter ||
                a.free_sampling_margin != b.free_sampling_margin || a.parallel != b.parallel || a.logodd_free != b.logodd_free || a.logodd_occupied != b.logodd_occupied ||
                a.logodd_variance != b.logodd_variance) {
                return false;
            }
            if (m_sp_gp_ != other.m_sp_gp_) { return false; }
            return m_map_boundary_.center == other.m_map_boundary_.center && m_map_boundary_.half_sizes == other.m_map_boundary_.half_sizes;
        }

        [[nodiscard]] bool
        operator!=(const SpGpOccupancyMap &other) const {
            return !(*this == other);
        }

        [[nodiscard]] bool
        Write(std::ostream &s) const {  // :164-203
            namespace ser = b200::serialization;
            return ser::WriteTokens(s, {{"setting",
                                         [this](std::ostream &o) {
                                             const Setting &g = *m_setting_;
                                             o.precision(17);
                                             o << g.min_distance << ' ' << g.max_distance << ' ' << g.free_points_per_meter << ' ' << g.free_sampling_margin << ' ' << g.parallel << ' '
                                               << g.logodd_free << ' ' << g.logodd_occupied << ' ' << g.logodd_variance;
                                             return o.good();
                                         }},
                                        {"sp_gp", [this](std::ostream &o) { return m_sp_gp_.Write(o); }},
                                        {"map_boundary", [this](std::ostream &o) { return ser::SaveMatrix(o, m_map_boundary_.center) && ser::SaveMatrix(o, m_map_boundary_.half_sizes); }},
                                        {"generator", [this](std::ostream &o) { o << m_generator_; return o.good(); }}});
        }

        [[nodiscard]] bool
        Read(std::istream &s) {  // :205-242
            namespace ser = b200::serialization;
            return ser::ReadTokens(s, {{"setting",
                                        [this](std::istream &i) {
                                            Setting &g = *m_setting_;
                                            i >> g.min_distance >> g.max_distance >> g.free_points_per_meter >> g.free_sampling_margin >> g.parallel >> g.logodd_free >> g.logodd_occupied >>
                                                g.logodd_variance;
                                            return !i.fail();
                                        }},
                                       {"sp_gp", [this](std::istream &i) { return m_sp_gp_.Read(i); }},
                                       {"map_boundary", [this](std::istream &i) { return ser::LoadVector(i, m_map_boundary_.center) && ser::LoadVector(i, m_map_boundary_.half_sizes); }},
                                       {"generator", [this](std::istream &i) { i >> m_generator_; return !i.fail(); }}});
        }

    private:
        static std::shared_ptr<Setting>
        CheckedSetting(std::shared_ptr<Setting> setting) {
            b200::AssertM(setting != nullptr, "setting is nullptr.");
            return setting;
        }

        std::shared_ptr<Setting> m_setting_ = nullptr;
        SpGp m_sp_gp_;
        AabbD m_map_boundary_;
        std::mt19937_64 m_generator_;
    };

    using SpGpOccupancyMap2Dd = SpGpOccupancyMap<double, 2>;
    using SpGpOccupancyMap2Df = SpGpOccupancyMap<float, 2>;
    using SpGpOccupancyMap3Dd = SpGpOccupancyMap<double, 3>;
    using SpGpOccupancyMap3Df = SpGpOccupancyMap<float, 3>;

}  // namespace erl::gaussian_process
