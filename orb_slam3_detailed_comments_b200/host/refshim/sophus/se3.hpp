// refshim/sophus/se3.hpp -- stand-in for Sophus::SE3<T> (Thirdparty/Sophus/sophus/se3.hpp): the members the host translation
// units call, over refshim/Eigen.  Compile-check / marshaling tests only; a real integration uses the vendored Sophus.
#pragma once
#include "Eigen/Core"

namespace Sophus {

template <typename T>
class SE3 {
   public:
    SE3() {}
    SE3(const Eigen::Quaternion<T>& q, const Eigen::Matrix<T, 3, 1>& t) : q_(q), t_(t) { q_.normalize(); }   // SO3(quat) normalises
#if defined(ORB_REFSHIM_LIBA) || defined(ORB_REFSHIM_FUSE)
    SE3(const Eigen::Matrix<T, 3, 3>& R, const Eigen::Matrix<T, 3, 1>& t) : q_(Eigen::Quaternion<T>(R)), t_(t) { q_.normalize(); }
#endif
    const Eigen::Quaternion<T>& unit_quaternion() const { return q_; }
    const Eigen::Matrix<T, 3, 1>& translation() const { return t_; }
    Eigen::Matrix<T, 3, 1>& translation() { return t_; }
    Eigen::Matrix<T, 3, 3> rotationMatrix() const { return q_.toRotationMatrix(); }
    SE3 inverse() const {
        SE3 o;
        o.q_ = q_.conjugate();
        o.t_ = -(o.q_.toRotationMatrix() * t_);
        return o;
    }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return q_.toRotationMatrix() * p + t_; }
    SE3 operator*(const SE3& b) const {
        SE3 o;
        o.q_ = q_ * b.q_;
        o.t_ = q_.toRotationMatrix() * b.t_ + t_;
        return o;
    }
    template <typename U>
    SE3<U> cast() const { SE3<U> o(q_.template cast<U>(), t_.template cast<U>()); return o; }

   private:
    Eigen::Quaternion<T> q_;
    Eigen::Matrix<T, 3, 1> t_;
};
#ifdef ORB_REFSHIM_TRI
template <typename T>
struct SO3 {
    static Eigen::Matrix<T, 3, 3> hat(const Eigen::Matrix<T, 3, 1>& w) {   // so3.hpp: [0 -c b; c 0 -a; -b a 0]
        Eigen::Matrix<T, 3, 3> m;
        m(0, 1) = -w(2); m(0, 2) = w(1); m(1, 0) = w(2); m(1, 2) = -w(0); m(2, 0) = -w(1); m(2, 1) = w(0);
        return m;
    }
};
typedef SO3<float> SO3f;
#endif
typedef SE3<float> SE3f;
typedef SE3<double> SE3d;

}  // namespace Sophus
