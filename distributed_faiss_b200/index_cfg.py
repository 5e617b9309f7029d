"""Per-index configuration object, field-compatible with the reference's `IndexCfg`
(distributed_faiss/index_cfg.py:11-64): same constructor keywords and defaults, unknown
keywords collected in `.extra` (e.g. `code_size`, `bits_per_vector` for the knnlm builder,
index.py:44-45), JSON round trip, `get_metric()` returning the faiss enum VALUES
(METRIC_INNER_PRODUCT = 0, METRIC_L2 = 1) that libdfx also uses.
"""
import json

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


class IndexCfg:
    def __init__(self, index_builder_type: str = None, faiss_factory: str = None, dim: int = 768,
                 train_num: int = 0, train_ratio: int = 1.0, centroids: int = 0, metric: str = "dot",
                 nprobe: int = 1, infer_centroids=False, buffer_bsz: int = 50000,
                 save_interval_sec: int = -1, index_storage_dir: str = None,
                 custom_meta_id_idx: int = 0, **kwargs):
        self.index_builder_type = index_builder_type
        self.faiss_factory = faiss_factory
        self.dim = int(dim)  # JSON configs carry it as a string (tests/test_index_config.json)
        self.train_num = train_num
        self.train_ratio = train_ratio
        self.centroids = centroids
        self.metric = metric
        self.nprobe = nprobe
        self.infer_centroids = infer_centroids
        self.buffer_bsz = buffer_bsz
        self.save_interval_sec = save_interval_sec
        self.index_storage_dir = index_storage_dir
        self.custom_meta_id_idx = custom_meta_id_idx
        # a cfg.json written by to_json_string() nests the extras under "extra"; flatten them so
        # that code_size / bits_per_vector survive a save -> load round trip
        nested = kwargs.pop("extra", None)
        if isinstance(nested, dict):
            kwargs = {**nested, **kwargs}
        self.extra = kwargs

    def get_metric(self) -> int:
        try:
            return {"dot": METRIC_INNER_PRODUCT, "l2": METRIC_L2}[self.metric]
        except KeyError:
            raise RuntimeError("Only dot and l2 metrics are supported.")

    @classmethod
    def from_json(cls, json_path):
        with open(json_path, "r") as fh:
            return cls(**json.load(fh))

    def to_json_string(self) -> str:
        # `extra` is written as a nested object, like the reference does
        return json.dumps(self.__dict__, default=lambda o: o.__dict__, sort_keys=True, indent=4)

    def __repr__(self) -> str:
        return f"<IndexCFG: {self.__dict__}>"
