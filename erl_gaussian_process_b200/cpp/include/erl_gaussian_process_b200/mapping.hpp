// Mapping<Dtype> — mirror of include/erl_gaussian_process/mapping.hpp (same Setting fields, enum values and
// map / inv members); the device applies the same warps inside the gather and predict kernels.
#pragma once

#include <cmath>
#include <functional>
#include <limits>
#include <memory>
#include <stdexcept>

namespace erl::gaussian_process {

    enum class MappingType { kIdentity = 0, kInverse = 1, kInverseSqrt = 2, kExp = 3, kLog = 4, kTanh = 5, kSigmoid = 6, kUnknown = 7 };

    template<typename Dtype>
    class Mapping {
    public:
        struct Setting {
            MappingType type = MappingType::kUnknown;
            Dtype scale = 1.0;
        };

    protected:
        std::shared_ptr<Setting> m_setting_;

    public:
        std::function<Dtype(Dtype)> map;
        std::function<Dtype(Dtype)> inv;

        static std::shared_ptr<Mapping>
        Create() {
            return Create(std::make_shared<Setting>());
        }

        static std::shared_ptr<Mapping>
        Create(std::shared_ptr<Setting> setting) {
            return std::shared_ptr<Mapping>(new Mapping(std::move(setting)));
        }

        [[nodiscard]] std::shared_ptr<Setting>
        GetSetting() const {
            return m_setting_;
        }

    private:
        explicit Mapping(std::shared_ptr<Setting> setting)
            : m_setting_(std::move(setting)) {
            const Setting *s = m_setting_.get();
            switch (s->type) {  // src/mapping.cpp:112-164
                case MappingType::kIdentity:
                    map = [](const Dtype x) { return x; };
                    inv = map;
                    break;
                case MappingType::kInverse:
                    map = [](const Dtype x) { return Dtype(1) / x; };
                    inv = map;
                    break;
                case MappingType::kInverseSqrt:
                    map = [](const Dtype x) { return Dtype(1) / std::sqrt(x); };
                    inv = [](const Dtype y) { return Dtype(1) / (y * y); };
                    break;
                case MappingType::kExp:
                    map = [s](const Dtype x) { return std::exp(-s->scale * x); };
                    inv = [s](const Dtype y) { return -std::log(y) / s->scale; };
                    break;
                case MappingType::kLog:
                    map = [s](const Dtype x) { return std::log(s->scale * x); };
                    inv = [s](const Dtype y) { return std::exp(y) / s->scale; };
                    break;
                case MappingType::kTanh:
                    map = [s](const Dtype x) { return std::tanh(s->scale * x); };
                    inv = [s](const Dtype y) { return std::atanh(y) / s->scale; };
                    break;
                case MappingType::kSigmoid:
                    map = [s](const Dtype x) -> Dtype { return Dtype(1) / (Dtype(1) + std::exp(-s->scale * x)); };
                    inv = [s](const Dtype y) -> Dtype {
                        if (y >= Dtype(1)) { return std::numeric_limits<Dtype>::infinity() / s->scale; }
                        if (y <= Dtype(0)) { return -std::numeric_limits<Dtype>::infinity() / s->scale; }
                        return std::log(y / (Dtype(1) - y)) / s->scale;
                    };
                    break;
                case MappingType::kUnknown:
                default:
                    throw std::logic_error("Mapping type is kUnknown, which is unexpected.");
            }
        }
    };

    using MappingD = Mapping<double>;
    using MappingF = Mapping<float>;
}  // namespace erl::gaussian_process
