"""Global SCvx parameters -- same names and shipped values as SCvx/global_parameters.py:4-18."""

K = 100                 # number of discretisation points
MAX_ITER = 30           # maximum SCvx iterations
TRUST_RADIUS0 = 100.0   # initial trust-region radius
CONV_TOL = 1e-3         # convergence tolerance
WEIGHT_NU = 1e4         # defect (virtual control) penalty
WEIGHT_SLACK = 1e6      # obstacle slack penalty
WEIGHT_SIGMA = 100.0    # time-scale weight
