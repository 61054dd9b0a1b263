"""Episode monitor for single envs -- same files and info['episode'] records as the reference's
baselines/bench/monitor.py:10-163 (`<name>.monitor.csv`: a `# {json header}` line, then a csv with columns r,l,t
(+ extra keys)), so the reference's plotting tools (`results_plotter`, `plot_util.load_results`) read our logs.
No gym dependency: wraps any object with reset()/step()."""
import csv
import json
import os.path as osp
import time
from glob import glob

__all__ = ['Monitor', 'ResultsWriter', 'get_monitor_files', 'load_results']


class ResultsWriter:
    def __init__(self, filename, header='', extra_keys=()):
        assert filename is not None
        if not filename.endswith(Monitor.EXT):
            filename = osp.join(filename, Monitor.EXT) if osp.isdir(filename) else filename + "." + Monitor.EXT
        self.f = open(filename, "wt")
        if isinstance(header, dict):
            header = '# {} \n'.format(json.dumps(header))
        self.f.write(header)
        self.logger = csv.DictWriter(self.f, fieldnames=('r', 'l', 't') + tuple(extra_keys))
        self.logger.writeheader()
        self.f.flush()

    def write_row(self, epinfo):
        self.logger.writerow(epinfo)
        self.f.flush()


class Monitor:
    EXT = "monitor.csv"

    def __init__(self, env, filename, allow_early_resets=False, reset_keywords=(), info_keywords=()):
        self.env = env
        self.observation_space, self.action_space = env.observation_space, env.action_space
        self.tstart = time.time()
        spec = getattr(env, "spec", None)
        self.results_writer = ResultsWriter(
            filename, header={"t_start": self.tstart, "env_id": getattr(spec, "id", None)},
            extra_keys=tuple(reset_keywords) + tuple(info_keywords)) if filename else None
        self.reset_keywords, self.info_keywords = reset_keywords, info_keywords
        self.allow_early_resets = allow_early_resets
        self.rewards = None
        self.needs_reset = True
        self.episode_rewards, self.episode_lengths, self.episode_times = [], [], []
        self.total_steps = 0
        self.current_reset_info = {}

    def __getattr__(self, name):                      # gym.Wrapper behaviour: fall through to the wrapped env
        if name.startswith('_'):
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):
        if not self.allow_early_resets and not self.needs_reset:
            raise RuntimeError("Tried to reset an environment before done. If you want to allow early resets, "
                               "wrap your env with Monitor(env, path, allow_early_resets=True)")
        self.rewards = []
        self.needs_reset = False
        for k in self.reset_keywords:
            if kwargs.get(k) is None:
                raise ValueError('Expected you to pass kwarg %s into reset' % k)
            self.current_reset_info[k] = kwargs[k]
        return self.env.reset(**kwargs)

    def step(self, action):
        if self.needs_reset:
            raise RuntimeError("Tried to step environment that needs reset")
        ob, rew, done, info = self.env.step(action)
        self.rewards.append(rew)
        if done:
            self.needs_reset = True
            eprew, eplen = sum(self.rewards), len(self.rewards)
            now = time.time() - self.tstart
            epinfo = {"r": round(eprew, 6), "l": eplen, "t": round(now, 6)}
            epinfo.update({k: info[k] for k in self.info_keywords})
            self.episode_rewards.append(eprew)
            self.episode_lengths.append(eplen)
            self.episode_times.append(now)
            epinfo.update(self.current_reset_info)
            if self.results_writer:
                self.results_writer.write_row(epinfo)
            assert isinstance(info, dict)
            info['episode'] = epinfo
        self.total_steps += 1
        return ob, rew, done, info

    def close(self):
        if hasattr(self.env, "close"):
            self.env.close()
        if self.results_writer:
            self.results_writer.f.close()

    def get_total_steps(self):
        return self.total_steps

    def get_episode_rewards(self):
        return self.episode_rewards

    def get_episode_lengths(self):
        return self.episode_lengths

    def get_episode_times(self):
        return self.episode_times


def get_monitor_files(dir):
    return glob(osp.join(dir, "*" + Monitor.EXT))


def load_results(dir):
    """All episodes of every *monitor.csv under `dir`, sorted by absolute time; returns a pandas DataFrame with
    columns r,l,t (t relative to the earliest t_start) and a `.headers` attribute, like monitor.py:128-163."""
    import pandas
    files = get_monitor_files(dir)
    if not files:
        raise FileNotFoundError("no monitor files of the form *%s found in %s" % (Monitor.EXT, dir))
    dfs, headers = [], []
    for fname in files:
        with open(fname, 'rt') as fh:
            first = fh.readline()
            if not first:
                continue
            assert first[0] == '#'
            header = json.loads(first[1:])
            df = pandas.read_csv(fh, index_col=None)
            headers.append(header)
            df['t'] += header['t_start']
        dfs.append(df)
    df = pandas.concat(dfs)
    df.sort_values('t', inplace=True)
    df.reset_index(inplace=True)
    df['t'] -= min(h['t_start'] for h in headers)
    object.__setattr__(df, "headers", headers)          # plain attribute (pandas would warn about a new column)
    return df
