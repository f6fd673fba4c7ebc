// TEST INFRASTRUCTURE ONLY: the reference includes this header but uses nothing from it.
