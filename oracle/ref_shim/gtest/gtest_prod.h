// Test-only stand-in for gtest's gtest_prod.h (GoogleTest is not installed).
#ifndef TG_GTEST_PROD_SHIM
#define TG_GTEST_PROD_SHIM
#define FRIEND_TEST(test_case_name, test_name) friend class test_case_name##_##test_name##_Test
#endif
