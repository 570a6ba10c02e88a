"""CPU restatement of RL.py's reward-to-go and rollout loops -- ORACLE, test infrastructure only.

``rl_solve`` follows RL_AC.RL_Solve (RL.py:145-189) for ``env_RL = 0`` (every conf), pinned
against the reference run in the build container (tests/golden/rtg_*.npz).
``create_to_init`` follows RL_AC.create_TO_init (RL.py:197-233).
"""
import numpy as np


def rl_solve(conf, TO_states, TO_step_cost):
    """Returns (state_arr, partial_rtg, total_rtg, state_next_rollout, done, rwrd, term, ep_return)."""
    T = len(TO_step_cost) - 1                               # NSTEPS_SH
    rwrd = -np.asarray(TO_step_cost, dtype=float)           # RL.py:168
    state = np.asarray(TO_states, dtype=float)
    s_next = np.zeros((T + 1, conf.nb_state))
    partial = np.empty(T + 1)
    total = np.empty(T + 1)
    term = np.zeros(T + 1)
    term[-1] = 1
    done = np.zeros(T + 1)
    ep_return = sum(rwrd)
    for i in range(T + 1):                                  # RL.py:173-187
        if conf.MC:
            final = T
            done[i] = 1
        else:
            final = min(i + conf.nsteps_TD_N, T)
            if final == T:
                done[i] = 1
            else:
                s_next[i, :] = state[final + 1, :]
        partial[i] = np.float32(sum(rwrd[i:final + 1]))     # quirk Q12: float32 rounding
        total[i] = np.float32(sum(rwrd[i:T + 1]))
    return state, partial, total, s_next, done, rwrd, term, ep_return


def horizon(conf, t0):
    """NSTEPS_SH = NSTEPS - int(t0/dt)  (RL.py:201; fp64 division then truncation)."""
    return conf.NSTEPS - int(t0 / conf.dt)


def create_to_init(conf, env, actor_eval, ep, ICS):
    """actor_eval(state[1, ns] float64) -> action[na] (float32 values); ep == 0 -> zero controls.
    Returns (ICS, states[T+1, ns], controls[T, na], T, success)."""
    T = horizon(conf, ICS[-1])
    if T == 0:
        return None, None, None, None, 0
    controls = np.zeros((T, conf.nb_action))
    states = np.zeros((T + 1, conf.nb_state))
    states[0, :] = ICS
    for i in range(T):                                      # RL.py:223-231
        if ep == 0:
            controls[i, :] = np.zeros(conf.nb_action)
        else:
            controls[i, :] = actor_eval(np.array([states[i, :]]))
        states[i + 1, :] = env.simulate(states[i, :], controls[i, :])
        if np.isnan(states[i + 1, :]).any():
            return None, None, None, None, 0
    return ICS, states, controls, T, 1
