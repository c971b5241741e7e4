// Token-stream serialization of the drop-in classes: Write(std::ostream &) / Read(std::istream &) with the token names and the token
// order of the reference (src/vanilla_gp.cpp:606-790, src/noisy_input_gp.cpp:436-606, 951-1160, src/lidar_gp_2d.cpp:495-635,
// src/range_sensor_gp_3d.cpp:472-655).  The framing helpers the reference calls (common::WriteTokens / ReadTokens,
// SaveEigenMatrixToBinaryStream, Yamlable::Write) live in erl_common, whose sources are absent here, so the framing below is this
// library's own and byte compatibility with files written by the reference is NOT claimed:
//   token line:   <name>'\n'   payload   '\n'          (tokens are read back in the order they were written; an unknown or
//   terminator:   "end_of_tokens\n"                      missing token makes Read() return false)
//   scalar:       text (operator<<)
//   matrix:       int64 rows, int64 cols, rows * cols scalars, column-major, raw bytes
// What Read() restores is the HOST state (setting, flags, train set, K / L / alpha); the device state is rebuilt by replaying the
// training on the GPU, which is bit-reproducible (tests/test_gpu_dense.py::test_vanilla_train_is_deterministic), and Read() fails
// if the rebuilt L differs from the stored one.
#pragma once

#include <cstdint>
#include <functional>
#include <istream>
#include <ostream>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace erl::gaussian_process::b200::serialization {

    using WritePairs = std::vector<std::pair<std::string, std::function<bool(std::ostream &)>>>;
    using ReadPairs = std::vector<std::pair<std::string, std::function<bool(std::istream &)>>>;

    inline void
    SkipLine(std::istream &s) {
        std::string rest;
        std::getline(s, rest);
    }

    inline bool
    WriteTokens(std::ostream &s, const WritePairs &pairs) {
        for (const auto &[token, fn]: pairs) {
            s << token << '\n';
            if (!fn(s)) { return false; }
            s << '\n';
        }
        s << "end_of_tokens" << '\n';
        return s.good();
    }

    inline bool
    ReadTokens(std::istream &s, const ReadPairs &pairs) {
        std::string token;
        for (const auto &[expected, fn]: pairs) {
            if (!std::getline(s, token) || token != expected) { return false; }
            if (!fn(s)) { return false; }
            SkipLine(s);
        }
        return std::getline(s, token) && token == "end_of_tokens";
    }

    template<typename Matrix>
    inline bool
    SaveMatrix(std::ostream &s, const Matrix &m) {
        const std::int64_t rows = m.rows(), cols = m.cols();
        s.write(reinterpret_cast<const char *>(&rows), sizeof(rows));
        s.write(reinterpret_cast<const char *>(&cols), sizeof(cols));
        if (rows * cols > 0) { s.write(reinterpret_cast<const char *>(m.data()), static_cast<std::streamsize>(sizeof(*m.data()) * rows * cols)); }
        return s.good();
    }

    template<typename Matrix>
    inline bool
    LoadMatrix(std::istream &s, Matrix &m) {
        std::int64_t rows = 0, cols = 0;
        s.read(reinterpret_cast<char *>(&rows), sizeof(rows));
        s.read(reinterpret_cast<char *>(&cols), sizeof(cols));
        if (!s.good() || rows < 0 || cols < 0 || rows > (1 << 24) || cols > (1 << 24) || rows * cols > (std::int64_t(1) << 31)) { return false; }  // a damaged header must not drive the allocation
        m.resize(rows, cols);
        if (rows * cols > 0) { s.read(reinterpret_cast<char *>(m.data()), static_cast<std::streamsize>(sizeof(*m.data()) * rows * cols)); }
        return s.good();
    }

    template<typename Vector>
    inline bool
    LoadVector(std::istream &s, Vector &v) {
        std::int64_t rows = 0, cols = 0;
        s.read(reinterpret_cast<char *>(&rows), sizeof(rows));
        s.read(reinterpret_cast<char *>(&cols), sizeof(cols));
        if (!s.good() || rows < 0 || rows > (std::int64_t(1) << 31) || (rows > 0 && cols != 1)) { return false; }
        v.resize(rows);
        if (rows > 0) { s.read(reinterpret_cast<char *>(v.data()), static_cast<std::streamsize>(sizeof(*v.data()) * rows)); }
        return s.good();
    }

    template<typename T>
    inline std::function<bool(std::ostream &)>
    ScalarWriter(const T &v) {
        return [&v](std::ostream &s) {
            s << v;
            return s.good();
        };
    }

    template<typename T>
    inline std::function<bool(std::istream &)>
    ScalarReader(T &v) {
        return [&v](std::istream &s) {
            s >> v;
            return !s.fail();
        };
    }

    // top-left rows x cols of two column-major matrices, element-wise (the reference compares topLeftCorner blocks, e.g. src/vanilla_gp.cpp:581-597)
    template<typename Matrix>
    inline bool
    SameTopLeft(const Matrix &a, const Matrix &b, const long rows, const long cols) {
        if (a.rows() < rows || a.cols() < cols || b.rows() < rows || b.cols() < cols) { return false; }
        for (long c = 0; c < cols; ++c) {
            for (long r = 0; r < rows; ++r) {
                if (!(a(r, c) == b(r, c))) { return false; }
            }
        }
        return true;
    }

    // Covariance::Setting / kernel_type as text (the reference writes them as YAML through Yamlable::Write)
    template<typename GpSetting>
    inline bool
    WriteGpSetting(std::ostream &s, const GpSetting &g) {
        s << g.kernel_type << '\n' << g.kernel_setting_type << '\n' << g.kernel->x_dim << ' ';
        s.precision(17);
        s << g.kernel->scale << ' ' << g.kernel->scale_mix << ' ' << g.kernel->weights.size();
        for (const auto w: g.kernel->weights) { s << ' ' << w; }
        s << ' ' << g.max_num_samples;
        return s.good();
    }

    template<typename GpSetting>
    inline bool
    ReadGpSetting(std::istream &s, GpSetting &g) {
        std::size_t nw = 0;
        if (!std::getline(s, g.kernel_type) || !std::getline(s, g.kernel_setting_type)) { return false; }
        s >> g.kernel->x_dim >> g.kernel->scale >> g.kernel->scale_mix >> nw;
        g.kernel->weights.resize(nw);
        for (auto &w: g.kernel->weights) { s >> w; }
        s >> g.max_num_samples;
        return !s.fail();
    }

    template<typename GpSetting>
    inline bool
    SameGpSetting(const GpSetting &a, const GpSetting &b) {
        return a.kernel_type == b.kernel_type && a.kernel_setting_type == b.kernel_setting_type && a.max_num_samples == b.max_num_samples && a.kernel->x_dim == b.kernel->x_dim &&
               a.kernel->scale == b.kernel->scale && a.kernel->scale_mix == b.kernel->scale_mix && a.kernel->weights == b.kernel->weights;
    }

    // (index_left, index_right, coord_left, coord_right) tables: count line, then the raw tuples (src/lidar_gp_2d.cpp:527-546)
    template<typename Partitions>
    inline bool
    WritePartitions(std::ostream &s, const Partitions &parts) {
        s << parts.size() << '\n';
        for (const auto &[il, ir, cl, cr]: parts) {
            s.write(reinterpret_cast<const char *>(&il), sizeof(il));
            s.write(reinterpret_cast<const char *>(&ir), sizeof(ir));
            s.write(reinterpret_cast<const char *>(&cl), sizeof(cl));
            s.write(reinterpret_cast<const char *>(&cr), sizeof(cr));
        }
        return s.good();
    }

    template<typename Partitions>
    inline bool
    ReadPartitions(std::istream &s, Partitions &parts) {
        std::size_t n = 0;
        s >> n;
        SkipLine(s);
        if (s.fail() || n > (1u << 24)) { return false; }
        parts.resize(n);
        for (auto &[il, ir, cl, cr]: parts) {
            s.read(reinterpret_cast<char *>(&il), sizeof(il));
            s.read(reinterpret_cast<char *>(&ir), sizeof(ir));
            s.read(reinterpret_cast<char *>(&cl), sizeof(cl));
            s.read(reinterpret_cast<char *>(&cr), sizeof(cr));
        }
        return s.good();
    }

    // partition GPs as the reference's "gps" token frames them (count line, one has_gp byte per GP, src/lidar_gp_2d.cpp:514-524); per GP
    // the state a caller of GetGps() can see: trained flag, number of samples, L (n x n), alpha (n)
    template<typename GpView>
    inline bool
    WritePartitionGp(std::ostream &s, const GpView &g) {
        const char trained = g.IsTrained() ? 1 : 0;
        const std::int64_t n = g.GetNumTrainSamples();
        s.write(&trained, 1);
        s.write(reinterpret_cast<const char *>(&n), sizeof(n));
        const auto &l = g.GetCholeskyDecomposition();
        const auto &a = g.GetAlpha();
        for (std::int64_t c = 0; c < n; ++c) {
            for (std::int64_t r = 0; r < n; ++r) { s.write(reinterpret_cast<const char *>(&l(r, c)), sizeof(l(r, c))); }
        }
        for (std::int64_t r = 0; r < n; ++r) { s.write(reinterpret_cast<const char *>(&a[r]), sizeof(a[r])); }
        return s.good();
    }

    // reads one partition GP record and compares it with the (re-trained) view: the check behind Read()
    template<typename GpView>
    inline bool
    ReadAndComparePartitionGp(std::istream &s, const GpView &g, const bool compare) {
        char trained = 0;
        std::int64_t n = 0;
        s.read(&trained, 1);
        s.read(reinterpret_cast<char *>(&n), sizeof(n));
        if (!s.good() || n < 0 || n > (1 << 20)) { return false; }
        using Scalar = std::remove_cv_t<std::remove_reference_t<decltype(g.GetAlpha()[0])>>;
        std::vector<Scalar> l(static_cast<std::size_t>(n * n)), a(static_cast<std::size_t>(n));
        if (n > 0) {
            s.read(reinterpret_cast<char *>(l.data()), static_cast<std::streamsize>(sizeof(Scalar) * n * n));
            s.read(reinterpret_cast<char *>(a.data()), static_cast<std::streamsize>(sizeof(Scalar) * n));
        }
        if (!s.good()) { return false; }
        if (!compare) { return true; }
        if ((trained != 0) != g.IsTrained() || n != g.GetNumTrainSamples()) { return false; }
        for (std::int64_t c = 0; c < n; ++c) {
            for (std::int64_t r = 0; r < n; ++r) {
                if (!(g.GetCholeskyDecomposition()(r, c) == l[static_cast<std::size_t>(r + c * n)])) { return false; }
            }
            if (!(g.GetAlpha()[c] == a[static_cast<std::size_t>(c)])) { return false; }
        }
        return true;
    }

    template<typename GpView>
    inline bool
    SamePartitionGp(const GpView &a, const GpView &b) {
        if (a.IsTrained() != b.IsTrained() || a.GetNumTrainSamples() != b.GetNumTrainSamples()) { return false; }
        const long n = a.GetNumTrainSamples();
        for (long c = 0; c < n; ++c) {
            for (long r = c; r < n; ++r) {
                if (!(a.GetCholeskyDecomposition()(r, c) == b.GetCholeskyDecomposition()(r, c))) { return false; }
            }
            if (!(a.GetAlpha()[c] == b.GetAlpha()[c])) { return false; }
        }
        return true;
    }

}  // namespace erl::gaussian_process::b200::serialization
