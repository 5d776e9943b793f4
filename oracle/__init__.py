"""oracle/ - CPU restatement of the reference's self-play hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import or execute anything in this directory, and only as the checker / reported CPU
baseline - never as the thing shipped.  The product (custom-alphazero_b200/) must not
import it; tests/test_product_isolation.py enforces that.

Parity status: PINNED.  The reference (neuronest/custom-alphazero) ships no tests, so the
pin is a set of golden vectors produced by running the unmodified reference in the build
container (tests/golden/make_golden.py, outputs committed as tests/golden/*.json):
environment fingerprints over 3 500 random playouts in six board configurations, six
single-search root statistics, ten deterministic full games (uniform and hash evaluators,
gravity and non-gravity, 3x3 ... 9x9) and three np.random-seeded stochastic games.
tests/test_oracle_golden.py checks both restatements (ref_port.py, c/az_oracle.c) against
all of them.

Files
  ref_port.py     object-per-node Python/numpy restatement (same data structures and cost
                  profile as the reference: int8 array boards, eager child boards); used for
                  small cases and as bench.py's "port" CPU baseline
  evaluators.py   the fixed evaluators (uniform/zero, hash) by specification
  c/az_oracle.c   plain-C restatement (array boards, pointer tree) for the sizes the
                  Python port cannot finish in seconds; built by oracle/Makefile into
                  oracle/_build/libaz_oracle.so
  c_oracle.py     ctypes loader for the C restatement
  c/chess_oracle.c    chess: mailbox rules + the reference's MCTS over them.  PARITY UNPINNED against
                      python-chess (the reference's third-party rules engine, absent here); pinned on published perft
                      counts (tests/test_chess_oracle.py)
  chess_ref.py        the reference's own chess layer restated (array, planes, play + mirror, action list) + loaders
  c/chess_hostcheck.cpp  the DEVICE rules header compiled for the host so the CPU suite can compare it with the oracle
"""
