"""Host-side pipeline around the learner (CPU only): VecNormalize / VecFrameStack pinned to reference outputs,
Monitor / VecMonitor file formats, SubprocVecEnv == DummyVecEnv, built-in envs, command-line plumbing."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


class _Scripted:
    """VecEnv replaying recorded (obs, rews, news)."""

    def __init__(self, obs, rews, news, ob_space):
        from baselines_b200.common import spaces
        self.obs, self.rews, self.news = obs, rews, news
        self.num_envs = obs.shape[1]
        self.observation_space, self.action_space = ob_space, spaces.Discrete(2)
        self.t = 0

    def reset(self):
        self.t = 0
        return self.obs[0]

    def step_async(self, actions):
        pass

    def step_wait(self):
        self.t += 1
        return self.obs[self.t], self.rews[self.t - 1], self.news[self.t - 1], [{}] * self.num_envs

    def close(self):
        pass


def test_vec_normalize_matches_reference_bit_exact():
    from baselines_b200.common import spaces
    from baselines_b200.common.vec_env import VecNormalize
    g = np.load(os.path.join(GOLDEN, "vec_normalize_trace.npz"))
    venv = _Scripted(g["obs"], g["rews"], g["news"], spaces.Box(-np.inf, np.inf, g["obs"].shape[2:], np.float32))
    env = VecNormalize(venv)
    assert np.array_equal(env.reset(), g["norm_obs"][0])
    for t in range(g["rews"].shape[0]):
        o, r, d, _ = env.step(None)
        assert np.array_equal(o, g["norm_obs"][t + 1]), t
        assert np.array_equal(r, g["norm_rews"][t]), t
    assert np.array_equal(env.ob_rms.mean, g["ob_mean"]) and np.array_equal(env.ob_rms.var, g["ob_var"])
    assert env.ob_rms.count == float(g["ob_count"]) and np.array_equal(env.ret_rms.var, g["ret_var"])


def test_builtin_envs_semantics():
    from baselines_b200 import envs
    e = envs.make("DiscreteIdentity-v0")
    e.seed(0)
    ob = e.reset()
    tot = 0
    for t in range(100):
        ob2, rew, done, _ = e.step(ob)                   # echo the observation: always rewarded
        tot += rew
        ob = ob2
        assert done == (t == 99)
    assert tot == 100
    ob = e.reset()
    assert e.step((ob + 1) % 10)[1] == 0
    b = envs.make("BoxIdentity-v0")
    b.seed(1)
    ob = b.reset()
    assert b.step(ob)[1] == 0.0 and b.observation_space.shape == (1,)
    c = envs.make("CartPole-v0")
    c.seed(3)
    ob = c.reset()
    assert ob.shape == (4,) and ob.dtype == np.float32 and np.all(np.abs(ob) <= 0.05)
    steps, done = 0, False
    while not done:
        ob, rew, done, info = c.step(1)                  # always push right: the pole falls within a few dozen steps
        steps += 1
        assert rew == 1.0
    assert 5 <= steps <= 60 and (abs(ob[2]) > c.env.theta_threshold or abs(ob[0]) > 2.4)
    # alternate pushes keep it up long enough to hit the 200-step cap at least sometimes; the cap must end it
    c.reset()
    for t in range(1000):
        ob, _, done, info = c.step(int(ob[2] + 0.3 * ob[3] > 0))      # a stabilising PD rule
        if done:
            break
    assert t == 199 and info.get("TimeLimit.truncated")
    a = envs.make("SyntheticAtari-v0")
    assert a.reset().shape == (84, 84, 1) and a.step(0)[0].dtype == np.uint8


def test_monitor_csv_format_and_episode_info(tmp_path):
    from baselines_b200 import envs
    from baselines_b200.bench import Monitor, load_results
    env = Monitor(envs.make("DiscreteIdentity-v0", episode_len=5), str(tmp_path / "0.0"), allow_early_resets=False)
    env.seed(0)
    with pytest.raises(RuntimeError):
        env.step(0)                                      # monitor.py:50-51: step before reset
    ob = env.reset()
    infos = []
    for _ in range(5):
        ob, rew, done, info = env.step(ob)
        infos.append(info)
    assert done and infos[-1]["episode"]["r"] == 5 and infos[-1]["episode"]["l"] == 5 and "episode" not in infos[0]
    with pytest.raises(RuntimeError):
        env.step(0)                                      # needs reset after done
    env.reset()
    with pytest.raises(RuntimeError):
        env.reset()                                      # early reset not allowed
    env.close()
    lines = open(tmp_path / "0.0.monitor.csv").read().splitlines()
    assert lines[0].startswith("#") and set(json.loads(lines[0][1:])) == {"t_start", "env_id"}
    assert json.loads(lines[0][1:])["env_id"] == "DiscreteIdentity-v0"
    assert lines[1] == "r,l,t" and lines[2].split(",")[:2] == ["5", "5"]
    df = load_results(str(tmp_path))
    assert list(df["r"]) == [5] and list(df["l"]) == [5] and len(df.headers) == 1


def _cartpole_thunk(i):
    def f():
        from baselines_b200 import envs
        e = envs.make("CartPole-v0")
        e.seed(100 + i)
        return e
    return f


@pytest.mark.parametrize("in_series", [1, 2])
def test_subproc_vec_env_equals_dummy(in_series):
    from baselines_b200.common.vec_env import DummyVecEnv, SubprocVecEnv
    n = 4
    sub = SubprocVecEnv([_cartpole_thunk(i) for i in range(n)], in_series=in_series)
    dum = DummyVecEnv([_cartpole_thunk(i) for i in range(n)])
    try:
        assert sub.num_envs == n and sub.observation_space.shape == (4,)
        assert np.array_equal(sub.reset(), dum.reset())
        rng = np.random.RandomState(0)
        ndone = 0
        for t in range(120):
            a = rng.randint(0, 2, n)
            o1, r1, d1, i1 = sub.step(a)
            o2, r2, d2, i2 = dum.step(a)
            assert np.array_equal(o1, o2) and np.array_equal(r1, r2) and np.array_equal(d1, d2), t
            assert r1.dtype == np.float32 and d1.dtype == np.bool_ and len(i1) == n
            ndone += int(d1.sum())
        assert ndone > 0                                  # auto-reset was exercised
    finally:
        sub.close()
        dum.close()
    assert sub.closed


def test_vec_monitor_and_frame_stack_compose(tmp_path):
    from baselines_b200.common.cmd_util import make_vec_env
    from baselines_b200.common.vec_env import VecFrameStack, VecMonitor
    venv = VecMonitor(make_vec_env("DiscreteIdentity-v0", "identity", 3, seed=0, force_dummy=True,
                                   env_kwargs=dict(episode_len=4)), filename=str(tmp_path / "vec"))
    ob = venv.reset()
    eps = []
    for _ in range(8):
        ob, rew, done, infos = venv.step(ob)
        eps += [i["episode"] for i in infos if "episode" in i]
    assert len(eps) == 6 and all(e["l"] == 4 for e in eps)
    lines = open(tmp_path / "vec.monitor.csv").read().splitlines()
    assert lines[1] == "r,l,t" and len(lines) == 2 + 6
    atari = VecFrameStack(make_vec_env("SyntheticAtari-v0", "atari", 2, seed=1, force_dummy=True), 4)
    o = atari.reset()
    assert o.shape == (2, 84, 84, 4) and atari.observation_space.shape == (84, 84, 4) and not o[..., :3].any()
    o2, _, _, _ = atari.step(np.zeros(2, dtype=np.int64))
    assert np.array_equal(o2[..., 2], o[..., 3]) or atari.venv.buf_dones.any()


def test_command_line_plumbing(tmp_path, monkeypatch):
    from baselines_b200 import run
    from baselines_b200.common.cmd_util import common_arg_parser, parse_unknown_args
    from baselines_b200.common.vec_env import SubprocVecEnv, VecFrameStack
    args, unknown = common_arg_parser().parse_known_args(
        ["--alg=ppo2", "--env=CartPole-v0", "--num_timesteps=3e4", "--num_env", "2", "--nsteps=64", "--lr", "1e-3",
         "--network=mlp", "--value_network=copy", "--cliprange=lambda f: 0.2 * f"])
    assert args.alg == "ppo2" and args.num_timesteps == 3e4 and args.num_env == 2 and args.play is False
    assert parse_unknown_args(unknown) == {"nsteps": "64", "lr": "1e-3", "value_network": "copy",
                                           "cliprange": "lambda f: 0.2 * f"}
    kw = run.parse_cmdline_kwargs(unknown)
    assert kw["nsteps"] == 64 and kw["lr"] == 1e-3 and kw["value_network"] == "copy" and kw["cliprange"](0.5) == 0.1
    assert run.get_env_type(args) == ("classic_control", "CartPole-v0")
    a2, _ = common_arg_parser().parse_known_args(["--env=atari"])
    assert run.get_env_type(a2) == ("atari", "SyntheticAtari-v0")
    a3, _ = common_arg_parser().parse_known_args(["--env=Foo-v0", "--env_type=mujoco"])
    assert run.get_env_type(a3) == ("mujoco", "Foo-v0")
    assert run.get_default_network("atari") == "cnn" and run.get_default_network("mujoco") == "mlp"
    d = run.get_learn_function_defaults("ppo2", "atari")
    assert d["nsteps"] == 128 and d["nminibatches"] == 4 and d["lr"](1.0) == 2.5e-4        # ppo2/defaults.py:15-22
    assert run.get_learn_function_defaults("ppo2", "classic_control") == {}
    assert run.get_learn_function_defaults("deepq", "atari")["prioritized_replay"] is True
    assert run.get_learn_function("ppo2").__module__.endswith("ppo2.ppo2")
    env = run.build_env(args)
    try:
        assert isinstance(env, SubprocVecEnv) and env.num_envs == 2
    finally:
        env.close()
    a4, _ = common_arg_parser().parse_known_args(["--env=SyntheticAtari-v0", "--num_env=2"])
    env = run.build_env(a4)
    try:
        assert isinstance(env, VecFrameStack) and env.observation_space.shape == (84, 84, 4) and env.frame_stack_device
    finally:
        env.close()


def test_schedules_and_explained_variance_match_reference():
    from baselines_b200.common.misc_util import explained_variance
    from baselines_b200.common.schedules import ConstantSchedule, LinearSchedule, PiecewiseSchedule
    g = np.load(os.path.join(GOLDEN, "host_misc.npz"))
    lin = LinearSchedule(schedule_timesteps=int(0.1 * 100000), initial_p=1.0, final_p=0.02)    # deepq.py:228-230
    beta = LinearSchedule(100000, initial_p=0.4, final_p=1.0)                                  # deepq.py:222-226
    pw = PiecewiseSchedule([(0, 1.0), (100, 0.5), (10000, 0.1)], outside_value=0.05)
    for i, t in enumerate(g["ts"]):
        assert lin.value(int(t)) == g["lin"][i] and beta.value(int(t)) == g["beta"][i] and pw.value(int(t)) == g["pw"][i]
    assert ConstantSchedule(0.7).value(3) == float(g["const"])
    assert explained_variance(g["yp"], g["y"]) == g["ev"] and explained_variance(g["y"], g["y"]) == g["ev_perfect"]
    assert np.isnan(explained_variance(g["yp"], np.ones_like(g["y"]))) and np.isnan(g["ev_const"])


def test_subproc_vec_env_spawn_context():
    """'spawn' start method (the reference's default, subproc_vec_env.py:44): thunks travel through cloudpickle."""
    from baselines_b200.common.vec_env import SubprocVecEnv
    env = SubprocVecEnv([_cartpole_thunk(i) for i in range(2)], context='spawn')
    try:
        o = env.reset()
        assert o.shape == (2, 4) and np.all(np.abs(o) <= 0.05)
        o2, r, d, infos = env.step(np.array([0, 1]))
        assert o2.shape == (2, 4) and r.tolist() == [1.0, 1.0] and not d.any() and len(infos) == 2
    finally:
        env.close()


def test_build_env_for_deepq_is_a_single_env():
    """run.py:97-99: deepq gets ONE env; for atari-type ids the frame stack is built in (4 stacked frames)."""
    from baselines_b200 import run
    from baselines_b200.common.cmd_util import common_arg_parser
    a, _ = common_arg_parser().parse_known_args(["--alg=deepq", "--env=SyntheticAtari-v0", "--seed=0"])
    env = run.build_env(a)
    try:
        ob = env.reset()
        assert ob.shape == (84, 84, 4) and env.observation_space.shape == (84, 84, 4) and env.action_space.n == 6
        ob2, rew, done, info = env.step(3)
        assert ob2.shape == (84, 84, 4) and isinstance(rew, float) and isinstance(done, bool) and isinstance(info, dict)
        assert np.array_equal(ob2[..., 2], ob[..., 3]) or done
    finally:
        env.close()
    a, _ = common_arg_parser().parse_known_args(["--alg=deepq", "--env=CartPole-v0", "--seed=0"])
    env = run.build_env(a)
    ob = env.reset()
    assert ob.shape == (4,) and not hasattr(env, "num_envs")
    env.close()


def test_logger_output_formats_and_env_selection(tmp_path, monkeypatch, capsys):
    """baselines/logger.py:174-190 (make_output_format) and :372-395 (configure): log.txt / progress.json /
    progress.csv writers, $OPENAI_LOG_FORMAT / $OPENAI_LOGDIR selection, "-rank%03i" suffix for ranks > 0, csv
    header growth with back-filled rows (:107-127), %-8.3g human formatting and 30-character truncation (:31-70)."""
    import json
    from baselines_b200 import logger
    d = str(tmp_path / "a")
    monkeypatch.delenv("OPENAI_LOG_FORMAT", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    try:
        logger.configure(d)                                         # default: stdout,log,csv
        logger.logkv("b", 2)
        logger.logkv("a_very_long_key_name_that_exceeds_thirty_chars", 1.23456789)
        logger.logkv_mean("m", 1.0)
        logger.logkv_mean("m", 3.0)
        logger.dumpkvs()
        logger.logkv("b", 3)
        logger.logkv("c", np.float32(0.5))
        logger.dumpkvs()
        logger.log("hello", "world")
        assert sorted(os.listdir(d)) == ["log.txt", "progress.csv"]
        rows = open(os.path.join(d, "progress.csv")).read().splitlines()
        assert rows[0] == "a_very_long_key_name_that_exceeds_thirty_chars,b,m,c"
        assert rows[1] == "1.23456789,2,2.0," and rows[2] == ",3,,0.5"
        txt = open(os.path.join(d, "log.txt")).read()
        assert "| a_very_long_key_name_that_e... | 1.23     |" in txt and "hello world" in txt
        assert txt.splitlines()[1].startswith("-----")
        out = capsys.readouterr().out
        assert "| b " in out and "Logging to" in out
        # explicit formats + json
        d2 = str(tmp_path / "b")
        logger.configure(d2, format_strs=["json"])
        logger.logkv("x", np.float64(1.5))
        logger.logkv("n", 7)
        logger.dumpkvs()
        assert os.listdir(d2) == ["progress.json"]
        assert json.loads(open(os.path.join(d2, "progress.json")).read()) == {"n": 7, "x": 1.5}
        # environment selection and rank suffix
        d3 = str(tmp_path / "c")
        monkeypatch.setenv("OPENAI_LOGDIR", d3)
        monkeypatch.setenv("OPENAI_LOG_FORMAT", "csv,json")
        logger.configure()
        logger.logkv("k", 1)
        logger.dumpkvs()
        assert sorted(os.listdir(d3)) == ["progress.csv", "progress.json"] and logger.get_dir() == d3
        monkeypatch.setenv("RANK", "2")
        monkeypatch.delenv("OPENAI_LOG_FORMAT_MPI", raising=False)
        logger.configure()
        logger.logkv("k", 1)
        logger.dumpkvs()
        assert "log-rank002.txt" in os.listdir(d3)
        import pytest
        with pytest.raises(NotImplementedError):
            logger.configure(d3, format_strs=["tensorboard"])
        with pytest.raises(ValueError):
            logger.configure(d3, format_strs=["nope"])
    finally:
        monkeypatch.delenv("OPENAI_LOGDIR", raising=False)
        monkeypatch.delenv("RANK", raising=False)
        logger.configure(None)
