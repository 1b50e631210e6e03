/* stgb.cpp itself only holds main(); its call sequence is restated in gen_driver.cpp (TEST INFRASTRUCTURE) */
