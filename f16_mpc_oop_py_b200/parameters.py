"""Constants of the plant, mirroring the reference's parameters.py (same names, same units, same values) for the
part the hot path uses: time step, names, units, bounds and the MPC/observation index maps."""
import numpy as np

dt, time_start, time_end = 0.001, 0., 10.  # parameters.py:22
fi_flag = 1                                 # parameters.py:26   1 hifi, 0 lofi
stab_flag = 0                               # parameters.py:31   0 -> xcg 0.25, 1 -> xcg 0.35
XCG = {0: 0.25, 1: 0.35}                    # C/nlplant.c:34 (the two reference builds)

states = ['npos', 'epos', 'h', 'phi', 'theta', 'psi', 'V', 'alpha', 'beta', 'p', 'q', 'r', 'T', 'dh', 'da', 'dr',
          'lf2', 'lf1']                                                     # parameters.py:116
inputs = ['T', 'dh', 'da', 'dr']                                            # parameters.py:117
x_units = ['ft', 'ft', 'ft', 'rad', 'rad', 'rad', 'ft/s', 'rad', 'rad', 'rad/s', 'rad/s', 'rad/s', 'lb', 'deg', 'deg',
           'deg', 'deg', 'deg']
u_units = ['lb', 'deg', 'deg', 'deg']

# parameters.py:59-95,122-129 (the reference compares raw state values with these numbers, whatever their unit)
x_ub = [np.inf, np.inf, 100000, np.inf, np.inf, np.inf, 900, 90, 30, 300, 100, 50, 19000, 25, 21.5, 30, 25, np.inf]
x_lb = [-np.inf, -np.inf, 0, -np.inf, -np.inf, -np.inf, 0, -20., -30., -300, -100, -50, 1000, -25, -21.5, -30., 0.,
        -np.inf]
u_ub = [19000, 25, 21.5, 30]
u_lb = [1000, -25, -21.5, -30.]
udot_ub = [10000, 60, 80, 120]
udot_lb = [-10000, -60, -80, -120]

observed_states = ['h', 'phi', 'theta', 'alpha', 'beta', 'p', 'q', 'r', 'lf2', 'lf1']   # parameters.py:134
mpc_states = ['phi', 'theta', 'alpha', 'beta', 'p', 'q', 'r', 'lf1', 'lf2']             # parameters.py:135
mpc_inputs = ['dh', 'da', 'dr']                                                         # parameters.py:136
mpc_controlled_states = ['p', 'q', 'r']                                                 # parameters.py:137

obs_x_idx = [states.index(s) for s in observed_states]   # [2,3,4,7,8,9,10,11,16,17]
mpc_x_idx = [states.index(s) for s in mpc_states]        # [3,4,7,8,9,10,11,17,16]
mpc_u_idx = [inputs.index(s) for s in mpc_inputs]        # [1,2,3]

# rough trim of parameters.py:36-55,105 (ft, rad, lb, deg)
m2f = 3.28084
x0 = np.array([0., 0., 3048. * m2f, 0., 0., 0., 213.36 * m2f, 1.0721 * np.pi / 180, 0., 0., 0., 0.,
               2886.6468, -2.0385, -0.087577, -0.03877, 0.3986, -1.0721 * np.pi / 180 * 180 / np.pi])
u0 = np.copy(x0[12:16])
