// oracle/ref_shim/dbow/boost/serialization/serialization.hpp -- TEST INFRASTRUCTURE.
// Thirdparty/DBoW2's BowVector.h / FeatureVector.h befriend boost::serialization::access and name base_object<> inside member
// templates that oracle/_ref never instantiates (map saving is out of scope); these declarations let the headers parse.
#pragma once
namespace boost { namespace serialization {
class access;
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
}}  // namespace boost::serialization
