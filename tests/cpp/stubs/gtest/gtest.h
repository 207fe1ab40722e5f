// stub of <gtest/gtest.h>: TEST_F fixtures, the assertions the reference's tests use, RUN_ALL_TESTS with gtest's output lines
#pragma once
#include <cstdio>
#include <functional>
#include <sstream>
#include <string>
#include <vector>
namespace testing {
class Test {
 public:
  virtual ~Test() {}
  virtual void SetUp() {}
  virtual void TearDown() {}
  virtual void TestBody() = 0;
};
struct Registry {
  struct Entry {
    std::string suite, name;
    std::function<Test*()> make;
  };
  std::vector<Entry> tests;
  bool current_failed = false;
  static Registry& get() {
    static Registry r;
    return r;
  }
};
struct Registrar {
  Registrar(const char* suite, const char* name, std::function<Test*()> make) {
    Registry::get().tests.push_back({suite, name, make});
  }
};
// collects `<< message` parts and reports on destruction
class Failure {
 public:
  Failure(const char* file, int line, const std::string& what) { os_ << file << ':' << line << ": Failure\n" << what; }
  Failure(const Failure& o) { os_ << o.os_.str(); }
  ~Failure() {
    std::printf("%s\n", os_.str().c_str());
    Registry::get().current_failed = true;
  }
  template <class T>
  Failure& operator<<(const T& v) {
    os_ << v;
    return *this;
  }
 private:
  std::ostringstream os_;
};
struct Void {  // lets `return Void() = Failure(...) << ...;` end a void test body (gtest's AssertHelper trick)
  void operator=(const Failure&) const {}
};
inline void InitGoogleTest(int*, char**) {}
inline int RunAll() {
  Registry& r = Registry::get();
  int failed = 0;
  std::printf("[==========] Running %zu tests.\n", r.tests.size());
  for (auto& e : r.tests) {
    std::printf("[ RUN      ] %s.%s\n", e.suite.c_str(), e.name.c_str());
    r.current_failed = false;
    try {
      Test* t = e.make();
      t->SetUp();
      t->TestBody();
      t->TearDown();
      delete t;
    } catch (const std::exception& ex) {
      std::printf("unknown file: Failure\nC++ exception with description \"%s\" thrown in the test body.\n", ex.what());
      r.current_failed = true;
    }
    if (r.current_failed) ++failed;
    std::printf("[  %s ] %s.%s\n", r.current_failed ? "FAILED " : "     OK", e.suite.c_str(), e.name.c_str());
  }
  std::printf("[==========] %zu tests ran.\n[  PASSED  ] %zu tests.\n", r.tests.size(), r.tests.size() - failed);
  if (failed) std::printf("[  FAILED  ] %d tests.\n", failed);
  std::fflush(stdout);
  return failed ? 1 : 0;
}
template <class A, class B>
inline std::string eq_message(const char* ea, const char* eb, const A&, const B&) {
  return std::string("Expected equality of these values:\n  ") + ea + "\n  " + eb;
}
}  // namespace testing
#define TEST_F(fixture, name)                                                                                      \
  class fixture##_##name##_Test : public fixture {                                                                 \
   public:                                                                                                         \
    void TestBody() override;                                                                                      \
  };                                                                                                               \
  static ::testing::Registrar fixture##_##name##_registrar(#fixture, #name, [] { return new fixture##_##name##_Test; }); \
  void fixture##_##name##_Test::TestBody()
#define GTEST_STUB_FAIL_(what) ::testing::Failure(__FILE__, __LINE__, what)
#define ADD_FAILURE() GTEST_STUB_FAIL_("Failed")
#define EXPECT_TRUE(c) \
  if (c) {             \
  } else               \
    GTEST_STUB_FAIL_(std::string("Value of: " #c "\n  Actual: false\nExpected: true"))
#define EXPECT_FALSE(c) EXPECT_TRUE(!(c))
#define EXPECT_EQ(a, b) \
  if ((a) == (b)) {     \
  } else                \
    GTEST_STUB_FAIL_(::testing::eq_message(#a, #b, a, b))
#define ASSERT_TRUE(c) \
  if (c) {             \
  } else               \
    return ::testing::Void() = GTEST_STUB_FAIL_(std::string("Value of: " #c "\n  Actual: false\nExpected: true"))
#define ASSERT_EQ(a, b) \
  if ((a) == (b)) {     \
  } else                \
    return ::testing::Void() = GTEST_STUB_FAIL_(::testing::eq_message(#a, #b, a, b))
#define RUN_ALL_TESTS() ::testing::RunAll()
