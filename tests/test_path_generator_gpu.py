"""PathGenerator through the public call on the GPU (C-ABI), against the fixtures recorded from the unmodified
reference's PathGenerator under the contract of tests/parity_contract.py."""
import numpy as np
import pytest

import helpers
import path_problems


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(path_problems.ALL))
def test_generate_path_on_the_gpu(native_lib, name):
    import parity_contract
    from trajectory_generator_b200 import batch
    from trajectory_generator_b200.path_generator import PathGenerator
    from trajectory_generator_b200.problem import pack_problem
    s = helpers.load_golden("path_generator.json")["problems"][name]["solve"]
    d, cc, kw = path_problems.ALL[name](helpers.product_namespace())
    gen = PathGenerator(d)
    cps = gen.generate_path(cc, **kw)
    r = gen.last_result
    pp = pack_problem(d, cc, kw.get("objective_function_type", "minimal_velocity_path"),
                      kw.get("num_intervals_free_space"), path_mode="indirect" if kw.get("isIndirect") else "direct")

    def cons(x):
        return batch.evaluate_host(pp.spec, pp.par[None], np.asarray(x, dtype=float)[None])["c"][0]
    parity_contract.check(name, s, d * pp.layout.N, r["x"], r["status"], r["nit"], r["fun"], cons, pp.layout.meq)
    assert np.array_equal(cps, np.asarray(r["x"])[:d * pp.layout.N].reshape(d, pp.layout.N))
