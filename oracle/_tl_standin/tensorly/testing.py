import numpy as np

assert_array_equal = np.testing.assert_array_equal
assert_allclose = np.testing.assert_allclose
