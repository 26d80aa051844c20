# Smoke test of the override: the reference's own basic problem (test/basic.jl:4-49) through the unchanged OSQP.jl API.
using OSQP, SparseArrays, Test
P = sparse([11.0 0.0; 0.0 0.0]); q = [3.0; 4]
A = sparse([-1.0 0; 0 -1; -1 -3; 2 5; 3 4]); u = [0.0; 0; -15; 100; 80]; l = -Inf * ones(5)
m = OSQP.Model()
OSQP.setup!(m; P = P, q = q, A = A, l = l, u = u, rho = 0.1, adaptive_rho = false, eps_abs = 1e-9, eps_rel = 1e-9,
            check_termination = 1, verbose = false, max_iter = 4000)
r = OSQP.solve!(m)
@test isapprox(r.x, [0.0; 5.0], atol = 1e-5)
@test isapprox(r.y, [1.666666666666; 0.0; 1.3333333; 0.0; 0.0], atol = 1e-5)
@test isapprox(r.info.obj_val, 20.0, atol = 1e-5)
@test occursin("b200", OSQP.version())          # "0.6.2-b200": the engine answered, not the stock libosqp
OSQP.update!(m; q = [1.0; 1.0])
@test OSQP.solve!(m).info.status == :Solved
println("OSQP.jl is running on the B200 engine: ", OSQP.version())
