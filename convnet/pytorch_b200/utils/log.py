"""Run bookkeeping with the reference's utils/log.py interface: ``setup_logging``, ``ResultsLog``
(CSV via pandas; the bokeh HTML plots of the reference are optional extras and are skipped when bokeh is not
installed), ``save_checkpoint`` and ``export_args_namespace``."""
import json
import logging
import os
import shutil

import torch


def export_args_namespace(args, filename):
    with open(filename, 'w') as fp:
        json.dump({k: v for k, v in vars(args).items()}, fp, sort_keys=True, indent=4, default=str)


def setup_logging(log_file='log.txt', resume=False, dummy=False):
    """DEBUG to file, INFO to console; ranks > 0 pass dummy=True and log nothing."""
    if dummy:
        logging.getLogger('dummy')
        return
    root = logging.getLogger()
    for h in list(root.handlers):
        root.removeHandler(h)
    root.setLevel(logging.DEBUG)
    mode = 'a' if (resume and os.path.isfile(log_file)) else 'w'
    to_file = logging.FileHandler(log_file, mode=mode)
    to_file.setLevel(logging.DEBUG)
    to_file.setFormatter(logging.Formatter('%(asctime)s - %(levelname)s - %(message)s', '%Y-%m-%d %H:%M:%S'))
    console = logging.StreamHandler()
    console.setLevel(logging.INFO)
    console.setFormatter(logging.Formatter('%(message)s'))
    root.addHandler(to_file)
    root.addHandler(console)


class ResultsLog(object):
    """Row-per-epoch results table saved as ``<path>.csv`` (and ``<path>.json`` for the params)."""

    def __init__(self, path='', title='', params=None, resume=False, data_format='csv'):
        import pandas as pd
        self._pd = pd
        self.data_path = '%s.%s' % (path, data_format)
        self.title = title
        self.rows = []
        self.results = pd.DataFrame()
        if params is not None:
            export_args_namespace(params, '%s.json' % path)
        if resume and os.path.isfile(self.data_path):
            self.load(self.data_path)

    def add(self, **kwargs):
        self.rows.append(kwargs)
        self.results = self._pd.DataFrame(self.rows)

    def load(self, path=None):
        path = path or self.data_path
        if os.path.isfile(path):
            self.results = self._pd.read_csv(path)
            self.rows = self.results.to_dict('records')
        else:
            raise ValueError('%s is not a file' % path)

    def save(self, title=None):
        self.results.to_csv(self.data_path, index=False, index_label=False)

    def plot(self, *args, **kwargs):
        """Plotting (bokeh in the reference) is not part of the training path; accepted and ignored."""
        return None

    def image(self, *args, **kwargs):
        return None

    def end(self):
        return None


def save_checkpoint(state, is_best, path='.', filename='checkpoint.pth.tar', save_all=False):
    target = os.path.join(path, filename)
    torch.save(state, target)
    if is_best:
        shutil.copyfile(target, os.path.join(path, 'model_best.pth.tar'))
    if save_all:
        shutil.copyfile(target, os.path.join(path, 'checkpoint_epoch_%s.pth.tar' % state['epoch']))
