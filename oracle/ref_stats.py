"""Random-policy statistics of the UNMODIFIED reference envs (run in the build container only; TEST INFRASTRUCTURE).

    python oracle/ref_stats.py penalty 12000 ; python oracle/ref_stats.py nopenalty 12000

Episode length, game score and invalid-move fraction over many episodes: the numbers tests/test_gpu_parity.py compares
the Philox mode with (it cannot follow MT19937 draw by draw).  An invalid move = Game2048.move(a, trial=True) says so
(nopenalty flavour) / the board comes back unchanged (penalty flavour: a valid move always spawns a tile).
Results of the committed numbers (12,000 episodes each, seeds below):
  penalty  : length 142.064 +- 0.427, score 1099.69 +- 4.87, invalid fraction 0.16395
  nopenalty: length 133.392 +- 0.390, score 1017.15 +- 4.50, invalid fraction 0.15393
"""
import sys, numpy as np, time, json
import os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim
flavour = sys.argv[1]; episodes = int(sys.argv[2])
mod = ref_shim.load_penalty_env() if flavour == "penalty" else ref_shim.load_nopenalty_env()
np.random.seed(12345 if flavour == "penalty" else 54321)
rs = np.random.RandomState(777)
env = mod.Game2048_env()
lengths, scores, invalid, steps = [], [], 0, 0
orig_move = mod.Game2048.move
state = {"valid": None, "depth": 0}
t0 = time.time()
for ep in range(episodes):
    env.reset()
    done, n = False, 0
    while not done:
        a = int(rs.randint(0, 4))
        before = np.array(env.game.board).copy()
        legal = None
        if flavour != "penalty":
            legal, _ = env.game.move(a, trial=True)      # no draws, no writes (Game2048_nopenalty_env.py:53-66)
        board, reward, done, mx = env.step(a)
        if flavour != "penalty":
            env.game.board = board
        after = np.array(board)
        # an invalid move leaves the board unchanged (no spawn)
        if (legal is False) or (legal is None and np.array_equal(before, after)): invalid += 1
        n += 1
    lengths.append(n); scores.append(int(env.score)); steps += n
L = np.array(lengths, float); S = np.array(scores, float)
out = {"flavour": flavour, "episodes": episodes, "steps": steps, "length_mean": L.mean(), "length_sem": L.std(ddof=1)/np.sqrt(episodes),
       "score_mean": S.mean(), "score_sem": S.std(ddof=1)/np.sqrt(episodes), "invalid_fraction": invalid/steps, "seconds": time.time()-t0}
print(json.dumps(out))
