// refshim/sophus/sim3.hpp -- stand-in for Sophus::Sim3<T> (see se3.hpp in this directory).
#pragma once
#include "sophus/se3.hpp"

namespace Sophus {

template <typename T>
class Sim3 {
   public:
    Sim3() : s_(1) {}
    Sim3(T s, const Eigen::Quaternion<T>& q, const Eigen::Matrix<T, 3, 1>& t) : s_(s), q_(q), t_(t) {}
    T scale() const { return s_; }
    Eigen::Matrix<T, 3, 3> rotationMatrix() const { return q_.toRotationMatrix(); }
    const Eigen::Quaternion<T>& quaternion() const { return q_; }
    const Eigen::Matrix<T, 3, 1>& translation() const { return t_; }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return (q_.toRotationMatrix() * p) * s_ + t_; }

   private:
    T s_;
    Eigen::Quaternion<T> q_;
    Eigen::Matrix<T, 3, 1> t_;
};
typedef Sim3<float> Sim3f;

}  // namespace Sophus
