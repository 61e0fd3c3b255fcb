"""dataLoader -- reads the reference's Kaldi-derived training files, same class and method names as
/root/reference/ctc_fast/dataLoader.py:6-95.

    feats%d.bin : float32 rows of `rawsize` features, all utterances of the file concatenated
    keys%d.txt  : "<utterance-key> <num_frames>" per line
    alis%d.txt  : "<utterance-key> <label> <label> ..." per line

loadDataFileDict returns (data_dict, alis, keys, sizes) with data_dict[key] an (imgsize x T) float32
array whose frames are contiguous (Fortran order), exactly what the reference hands to SGD.run.
The asynchronous prefetch uses a thread instead of the reference's forked child (dataLoader.py:22-36):
forking after CUDA initialisation is unsafe, and the loader is pure NumPy file IO.
"""
import os
import threading

import numpy as np


class DataLoader:
    def __init__(self, filedir_feat, rawsize, imgsize, filedir_ali=None, load_ali=True, load_data=True):
        self.filedir_feat = filedir_feat
        self.rawsize = rawsize
        self.imgsize = imgsize
        self.filedir_ali = filedir_feat if filedir_ali is None else filedir_ali
        self.load_ali = load_ali
        self.load_data = load_data
        self.p = None
        self._result = None

    def getDataAsynch(self):
        assert self.p is not None, "Error in order of asynch calls."
        self.p.join()
        self.p = None
        if isinstance(self._result, BaseException):
            raise self._result
        return self._result

    def loadDataFileAsynch(self, filenum):
        def work():
            try:
                self._result = self.loadDataFileDict(filenum)
            except BaseException as e:      # surfaced by getDataAsynch
                self._result = e
        self.p = threading.Thread(target=work)
        self.p.start()

    def loadDataFile(self, filenum):
        keyfile = os.path.join(self.filedir_feat, 'keys%d.txt' % filenum)
        alisfile = os.path.join(self.filedir_ali, 'alis%d.txt' % filenum)
        datafile = os.path.join(self.filedir_feat, 'feats%d.bin' % filenum)
        keys = sizes = data = None
        alis = []
        if self.load_ali:
            with open(alisfile, 'r') as fid:
                for l in fid.readlines():
                    l = l.split()
                    alis.append((l[0], l[1:]))
            alis = dict(alis)
        if self.load_data:
            if os.path.exists(keyfile):
                with open(keyfile, 'r') as keyf:
                    uttdat = [u.split() for u in keyf.readlines()]
                sizes = np.array([np.int32(u[1]) for u in uttdat])
                keys = [u[0] for u in uttdat]
            left = (self.rawsize - self.imgsize) // 2       # centre crop of the context window
            right = left + self.imgsize
            data = np.fromfile(datafile, np.float32).reshape(-1, self.rawsize)
            data = data[:np.sum(sizes), left:right]
            return data.T, alis, keys, sizes
        keys = list(alis.keys())
        return data, alis, keys, sizes

    def loadDataFileDict(self, filenum):
        data_mat, alis, keys, sizes = self.loadDataFile(filenum)
        if self.load_data:
            data_dict = {}
            startInd = 0
            for k, s in zip(keys, sizes):
                endInd = startInd + s
                data_dict[k] = np.copy(data_mat[:, startInd:endInd])
                startInd = endInd
            assert startInd == data_mat.shape[1]
            return data_dict, alis, keys, sizes
        return None, alis, keys, sizes


def write_synthetic_file(dirname, filenum, num_utts, rawsize, outputDim, T_range=(150, 250), L_range=(20, 40),
                         seed=33):
    """Write one synthetic file triple in the reference's on-disk format (for tests and examples)."""
    rng = np.random.RandomState(seed + filenum)
    os.makedirs(dirname, exist_ok=True)
    feats, keys, alis = [], [], []
    for u in range(num_utts):
        T = int(rng.randint(T_range[0], T_range[1] + 1))
        L = int(min(T, rng.randint(L_range[0], L_range[1] + 1)))
        feats.append(rng.randn(T, rawsize).astype(np.float32))
        key = "utt%d_%04d" % (filenum, u)
        keys.append("%s %d" % (key, T))
        alis.append(key + " " + " ".join(str(int(x)) for x in 1 + rng.randint(0, outputDim - 1, size=L)))
    np.concatenate(feats, axis=0).tofile(os.path.join(dirname, 'feats%d.bin' % filenum))
    open(os.path.join(dirname, 'keys%d.txt' % filenum), 'w').write("\n".join(keys) + "\n")
    open(os.path.join(dirname, 'alis%d.txt' % filenum), 'w').write("\n".join(alis) + "\n")
