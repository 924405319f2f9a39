// Stand-in for nanobind (absent from this image): the module definition at the end of the
// reference's evaluator.cpp compiles into a function nobody calls; the C API in
// oracle/ref_evaluator_capi.cpp drives the reference's classes directly.
#pragma once
#include <cstddef>

namespace nanobind {
struct module_ {
  template <typename... A>
  module_ &def(A &&...) { return *this; }
};
template <typename... A>
struct init {};
struct arg {
  explicit arg(const char *) {}
  template <typename T>
  arg &operator=(T &&) { return *this; }
};
template <typename T>
struct class_ {
  template <typename... A>
  class_(A &&...) {}
  template <typename... A>
  class_ &def(A &&...) { return *this; }
};
}  // namespace nanobind
#define NB_MODULE(name, var) [[maybe_unused]] static void nb_module_##name(nanobind::module_ &var)
