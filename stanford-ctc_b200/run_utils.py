"""run_utils -- the helpers runNNet.py needs, same names as /root/reference/ctc_fast/run_utils.py:10-91
(dump_config, load_config, CfgStruct, get_git_revision, get_hostname, touch_file, TimeString)."""
import datetime
import json
import os
import re
import subprocess


def dump_config(cfg, fname):
    json.dump(cfg, open(fname, 'w'), sort_keys=True, indent=4, separators=(',', ':'))


def load_config(fname):
    return json.load(open(fname, 'r'))


class CfgStruct:
    def __init__(self, **entries):
        self.__dict__.update(entries)


def get_git_revision():
    try:
        return subprocess.Popen(['git', 'rev-parse', '--short', 'HEAD'], stdout=subprocess.PIPE,
                                stderr=subprocess.DEVNULL).communicate()[0].decode().strip()
    except Exception:
        return ''


def get_hostname():
    import socket
    return socket.gethostname().split('.')[0]


def touch_file(fname):
    try:
        os.utime(fname, None)
    except Exception:
        open(fname, 'a').close()


class TimeString(object):
    def __str__(self):
        s = str(datetime.datetime.today())
        return s.split('.')[0].replace(' ', '').replace('-', '').replace(':', '')

    @classmethod
    def match(cls, s):
        return re.match(r'\d{14}$', s)
