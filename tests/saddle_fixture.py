"""The RNG-free CNOT fixture of the reference's test/test_lbfgsb_saddle_point.jl:18-112: two qubits with a static
sigma_y (x) sigma_y interaction, six single-qubit drives (ShapedAmplitude with a box shape == 1 on [0, T]),
guess E0 = 0.1, tlist = 0:0.001:1, four basis trajectories, J_T_sm."""
import numpy as np

from grape.jl_b200.optimize import hamiltonian, Trajectory, ShapedAmplitude, Control


def cnot_trajectories(E0=0.1, T=1.0, dt=0.001):
    I2 = np.eye(2, dtype=complex)
    sz = np.diag([1.0, -1.0]).astype(complex)
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sy = np.array([[0, -1j], [1j, 0]], dtype=complex)
    ops = [np.kron(sx, I2), np.kron(sy, I2), np.kron(sz, I2), np.kron(I2, sx), np.kron(I2, sy), np.kron(I2, sz)]
    H0 = np.pi / 2 * np.kron(sy, sy)
    tlist = np.arange(0.0, T + dt / 2, dt)                       # collect(range(0, T, step = dt))
    box = lambda t: 1.0 if 0.0 <= t <= T else 0.0                # QuantumControl.Shapes.box(t, 0, T)
    amps = [ShapedAmplitude(Control(lambda t: E0), box) for _ in ops]
    H = hamiltonian(H0, *zip(ops, amps))
    CNOT = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]], dtype=complex)
    basis = [np.eye(4, dtype=complex)[i] for i in range(4)]
    tgts = [CNOT.T @ b for b in basis]
    return [Trajectory(b, H, target_state=t) for b, t in zip(basis, tgts)], tlist
