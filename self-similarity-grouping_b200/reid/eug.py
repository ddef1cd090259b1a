"""Label-estimation half of reid/eug.py (EUG one-shot progressive labelling) on the CUDA path.

Kept from the reference (same names / signatures / return types): ``EUG(...)`` constructor, ``get_feature``
(eug.py:162-166), ``get_Dissimilarity_result`` (193-253, both the L2 and the re-ranking branch),
``estimate_label`` (255-271), ``select_top_data`` / ``select_top_true_data`` (275-288),
``generate_new_train_data`` (291-308).  The training half (``train``, ``resume``, data loaders built from the
reference's ``reid.utils.data``) is outside the pseudo-label hot path: those methods need the
reference's own packages on the path, or a ``loader_factory`` passed by the caller.
"""
import numpy as np
import torch

from .evaluators import extract_features
from .rerank_initial import re_ranking_init


def _first_min(dist):
    """Row minimum and the LOWEST index attaining it (np.argmin's tie rule, eug.py:208,231; torch's CUDA min/argmin
    do not promise which of several equal entries they return)."""
    dmin = dist.min(dim=1).values
    cols = torch.arange(dist.shape[1], device=dist.device).expand_as(dist)
    big = torch.full_like(cols, dist.shape[1])
    return dmin, torch.where(dist == dmin[:, None], cols, big).min(dim=1).values


class EUG():
    def __init__(self, model_name, batch_size, mode, num_classes, data_dir, l_data, u_data, save_path, print_freq,
                 dropout=0.5, pretrained_model=None, triplet=False, rerank=False, loader_factory=None):
        self.model_name = model_name
        self.num_classes = num_classes
        self.mode = mode
        self.data_dir = data_dir
        self.save_path = save_path
        self.l_data = [[f, l, 1.0] for f, l, _ in l_data]
        self.u_data = u_data
        self.l_label = np.array([label for _, label, _ in l_data])
        self.u_label = np.array([label for _, label, _ in u_data])
        self.batch_size = batch_size
        self.data_height, self.data_width, self.data_workers = 256, 128, 6
        self.eval_bs = batch_size
        self.dropout = dropout
        self.model = pretrained_model
        self.print_freq = print_freq
        self.num_instances = 4
        self.rerank = rerank
        self.loader_factory = loader_factory

    # ---- data / training: the reference's own packages (out of scope here)
    def get_dataloader(self, dataset, training=False):
        if self.loader_factory is not None:
            return self.loader_factory(dataset, training)
        try:
            from torch.utils.data import DataLoader
            from reid.utils.data import transforms as T
            from reid.utils.data.preprocessor import Preprocessor
        except ImportError:
            raise NotImplementedError("EUG.get_dataloader needs the reference's reid.utils.data on the path "
                                      "or a loader_factory(dataset, training)")
        normalizer = T.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
        tf = T.Compose([T.RectScale(self.data_height, self.data_width), T.ToTensor(), normalizer])
        return DataLoader(Preprocessor(dataset, root=self.data_dir, transform=tf), batch_size=self.eval_bs,
                          num_workers=self.data_workers, shuffle=False, pin_memory=True)

    def train(self, *args, **kwargs):
        raise NotImplementedError("EUG.train (fine-tuning) is outside the pseudo-label hot path (SURVEY.md 8f row f1)")

    def resume(self, ckpt_file, step):
        raise NotImplementedError("EUG.resume is outside the pseudo-label hot path")

    # ---- label estimation (hot-path API)
    def get_feature(self, dataset):
        """eug.py:162-166: eval-mode features (flip TTA, concatenated banks, one L2 norm) as an ndarray."""
        dataloader = self.get_dataloader(dataset, training=False)
        features, _ = extract_features(self.model, dataloader)
        return np.array([logit.numpy() for logit in features.values()])

    def get_Dissimilarity_result(self, weight=False):
        """eug.py:193-253 — nearest labelled neighbour by L2, or by re-ranked cosine distance (self.rerank)."""
        import torch
        from ssg_b200 import _lib
        from ssg_b200.rerank import sqdist, dot, re_ranking_init_blocks
        u_feas = self.get_feature(self.u_data)
        l_feas = self.get_feature(self.l_data)
        print("u_features", u_feas.shape, "l_features", l_feas.shape)
        dev = _lib.require_cuda()
        u = torch.from_numpy(np.ascontiguousarray(u_feas, np.float32)).to(dev)
        l = torch.from_numpy(np.ascontiguousarray(l_feas, np.float32)).to(dev)
        l_label = torch.from_numpy(self.l_label.astype(np.int64)).to(dev)
        confidence = None
        if not self.rerank:
            dist = sqdist(u, l, _lib.DIST_EXACT).sqrt()                     # np.linalg.norm(l_feas - u_fea, axis=1)
            dmin, index_min = _first_min(dist)
            scores = (-dmin).double().cpu().numpy()
        else:
            u_l, u_u, l_l = dot(u, l), dot(u, u), dot(l, l)                # np.dot blocks, eug.py:223-225 (ssg_dot)
            re_rank_dist = re_ranking_init_blocks(u_l, u_u, l_l)            # CUDA tensor [nu, nl]
            dmin, index_min = _first_min(re_rank_dist)
            scores = (-dmin).double().cpu().numpy()
            colmax = re_rank_dist.max(dim=0).values
            confidence = (1 - dmin / colmax[index_min]).double().cpu().numpy()   # eug.py:236
        labels = l_label[index_min].double().cpu().numpy()
        num_correct_pred = int((self.u_label == labels.astype(self.u_label.dtype)).sum())
        print("{} predictions on all the unlabeled data: {} of {} is correct, accuracy = {:0.3f}".format(
            self.mode, num_correct_pred, u_feas.shape[0], num_correct_pred / u_feas.shape[0]))
        if self.rerank and weight:
            return labels, scores, confidence
        return labels, scores

    def get_Classification_result(self):
        raise NotImplementedError("classification-mode EUG needs the classifier head training path (out of scope)")

    def estimate_label(self):
        print("label estimation by {} mode.".format(self.mode))
        if self.mode == "Dissimilarity":
            [pred_label, pred_score] = self.get_Dissimilarity_result()
            return pred_label, pred_score
        elif self.mode == "Classification":
            [pred_label, pred_score] = self.get_Classification_result()
            return pred_label, pred_score
        elif self.mode == 'Weight':
            [pred_label, pred_score, confidence] = self.get_Dissimilarity_result(True)
            return pred_label, pred_score, confidence
        else:
            raise ValueError

    def select_top_true_data(self, pred_label, pred_score, nums_to_select):
        v = np.zeros(len(pred_score))
        index = np.argsort(-pred_score)
        for i in range(nums_to_select):
            if pred_label[index[i]] != -1:
                v[index[i]] = 1
        return v.astype('bool')

    def select_top_data(self, pred_score, nums_to_select):
        v = np.zeros(len(pred_score))
        index = np.argsort(-pred_score)
        for i in range(nums_to_select):
            v[index[i]] = 1
        return v.astype('bool')

    def generate_new_train_data(self, sel_idx, pred_y):
        """ generate the next training data """
        seletcted_data = []
        correct, total = 0, 0
        for i, flag in enumerate(sel_idx):
            if flag:
                seletcted_data.append([self.u_data[i][0], int(pred_y[i]), self.u_data[i][2]])
                total += 1
                if self.u_label[i] == int(pred_y[i]):
                    correct += 1
        acc = correct / total
        new_train_data = self.l_data + seletcted_data
        print("selected pseudo-labeled data: {} of {} is correct, accuracy: {:0.4f}  new train data: {}".format(
            correct, len(seletcted_data), acc, len(new_train_data)))
        return new_train_data


def updata_lable(dataset, label, name, sample='random', load_path='random_split/', seed=0):
    raise NotImplementedError("updata_lable builds/pickles the one-shot dataset split (eug.py:325-381): dataset glue, "
                              "outside the pseudo-label hot path")
