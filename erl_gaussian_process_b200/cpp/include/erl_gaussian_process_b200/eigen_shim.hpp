// Minimal column-major Matrix / Vector / Ref subset with Eigen's spelling, used when Eigen is not on the
// include path (it is absent from this image).  With real Eigen on the include path (detected with __has_include,
// or forced with -DERL_GP_USE_EIGEN) the same headers build against it: only the subset below is used by the
// drop-in classes.  The shim lives in its own namespace and is aliased to `Eigen` only when Eigen itself is absent, so
// a translation unit can never see two definitions of Eigen::MatrixX.  -DERL_GP_NO_EIGEN forces the shim.
#pragma once

#include <cstdint>

#if !defined(ERL_GP_USE_EIGEN) && !defined(ERL_GP_NO_EIGEN) && defined(__has_include)
    #if __has_include(<Eigen/Dense>)
        #define ERL_GP_USE_EIGEN 1
    #endif
#endif

#ifdef ERL_GP_USE_EIGEN
    #include <Eigen/Dense>
namespace Eigen {
    // erl_common's typedefs (the reference's masks are Eigen matrices of bool)
    using VectorXb = Matrix<bool, Dynamic, 1>;
    using MatrixXb = Matrix<bool, Dynamic, Dynamic>;
}  // namespace Eigen
#else

    #include <cassert>
    #include <cstddef>
    #include <vector>

namespace erl_gp_eigen_shim {

    template<typename T>
    class MatrixX {
        std::vector<T> m_data_;
        long m_rows_ = 0, m_cols_ = 0;

    public:
        using Scalar = T;

        MatrixX() = default;

        MatrixX(const long rows, const long cols)
            : m_data_(static_cast<std::size_t>(rows * cols)),
              m_rows_(rows),
              m_cols_(cols) {}

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

        void
        resize(const long rows, const long cols) {  // like Eigen: contents unspecified after a size change
            m_data_.resize(static_cast<std::size_t>(rows * cols));
            m_rows_ = rows;
            m_cols_ = cols;
        }

        void
        setZero() {
            for (auto &v: m_data_) { v = T(0); }
        }

        void
        setConstant(const T value) {
            for (auto &v: m_data_) { v = value; }
        }

        T *
        data() {
            return m_data_.data();
        }

        const T *
        data() const {
            return m_data_.data();
        }

        T &
        operator()(const long r, const long c) {
            assert(r >= 0 && r < m_rows_ && c >= 0 && c < m_cols_);
            return m_data_[static_cast<std::size_t>(r + c * m_rows_)];
        }

        const T &
        operator()(const long r, const long c) const {
            assert(r >= 0 && r < m_rows_ && c >= 0 && c < m_cols_);
            return m_data_[static_cast<std::size_t>(r + c * m_rows_)];
        }

        [[nodiscard]] bool
        operator==(const MatrixX &other) const {
            return m_rows_ == other.m_rows_ && m_cols_ == other.m_cols_ && m_data_ == other.m_data_;
        }
    };

    template<typename T>
    class VectorX : public MatrixX<T> {
    public:
        VectorX() = default;

        explicit VectorX(const long n)
            : MatrixX<T>(n, 1) {}

        void
        resize(const long n) {
            MatrixX<T>::resize(n, 1);
        }

        T &
        operator[](const long i) {
            return (*this)(i, 0);
        }

        const T &
        operator[](const long i) const {
            return (*this)(i, 0);
        }
    };

    template<typename T>
    using Matrix2 = MatrixX<T>;  // 2 x 2, column-major
    template<typename T>
    using Matrix3 = MatrixX<T>;  // 3 x 3
    template<typename T>
    using Vector2 = VectorX<T>;
    template<typename T>
    using Vector3 = VectorX<T>;
    template<typename T>
    using Matrix3X = MatrixX<T>;  // 3 x n
    using VectorXb = VectorX<unsigned char>;
    using MatrixXb = MatrixX<unsigned char>;

    // `const Eigen::Ref<const MatrixX> &` collapses to `const MatrixX &`, `Eigen::Ref<VectorX>` to `VectorX &`
    template<typename M>
    using Ref = M &;

}  // namespace erl_gp_eigen_shim

namespace Eigen = erl_gp_eigen_shim;
#endif

namespace erl::gaussian_process::b200 {
    // The C ABI takes masks as uint8_t*.  The reference's masks are Eigen::VectorX<bool> (real Eigen) or the shim's
    // VectorX<unsigned char>: both are one byte per element holding 0 / 1.
    template<typename MaskVector>
    inline std::uint8_t *
    MaskData(MaskVector &v) {
        static_assert(sizeof(*v.data()) == 1, "mask elements must be one byte");
        return reinterpret_cast<std::uint8_t *>(v.data());
    }

    template<typename MaskVector>
    inline const std::uint8_t *
    MaskData(const MaskVector &v) {
        static_assert(sizeof(*v.data()) == 1, "mask elements must be one byte");
        return reinterpret_cast<const std::uint8_t *>(v.data());
    }
}  // namespace erl::gaussian_process::b200
