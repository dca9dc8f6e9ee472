"""Runtime knobs of the NRMS hot path.

Drop-in contract: the attribute NAMES and DEFAULT VALUES below are the ones the reference's
`Config` exposes (reference config.py:8-60 and `__nrms__` config.py:65-88), because its
scripts mutate them by name (`config.batch_size = 256`, `config.mode = 'demo'`,
`config.word_embedding_pretrained = ...`: run_demo.py:27-32, run_v0.py:44-50).  They are kept
in two tables here instead of an attribute-by-attribute constructor; knobs the NRMS path
never reads (BERT/entity/category sizes, CNN filter sizes) are carried as inert values so
that code which copies or prints them keeps working.
"""
from __future__ import annotations

import torch

# knobs the hot path actually reads -----------------------------------------------------------
_PATH_KNOBS = dict(
    n_words_title=20,          # T: words per title
    history_len=50,            # H: clicked-news slots per user
    sample_size=5,             # K negatives per positive (train batches carry K+1 slots)
    max_candidate_size=300,    # padded candidate slots of an eval impression
    word_embed_size=300,       # D
    dropout=0.2,
    batch_size=512,
    learning_rate=1e-3,
    num_epochs=5,
    random_seed=1998,
    mode='large',              # 'large' | 'demo': which title dictionary the loader opens
    data_path='./data_processed/',
    word_embedding_pretrained='all_word_embedding_v3.npz',
)

# knobs of the surrounding scripts / other model variants (inert for the nrms_v0 path; the `nrms` sibling
# plugin reads bert_embedding_pretrained here and bert_embed_size, news_feature_size, user_heads_num,
# query_vector_dim_large from __nrms__) --------------------------------------------------------
_INERT_KNOBS = dict(
    bert_embedding_pretrained='news_embeds_512.npz',
    entity_embedding_pretrained='entitiy_embeds.npz',
    save_path='./save_model/',
    train_data='train_datas.pkl', dev_data='dev_datas.pkl', test_data='test_datas.pkl',
    n_words_abst=40, save_flag=True, word_freq_threshold=3,
    category_nums=19, n_words=45800, subcategory_nums=294, cate_embed_size=100, entity_nums=10,
    eval_step=5000, require_improvement=10000, warm_up_steps=500, warm_up=False,
)

# what `__nrms__()` adds ------------------------------------------------------------------------
_NRMS_KNOBS = dict(
    query_vector_dim=200,      # Q: additive-attention width (read by the path)
    num_attention_heads=10,    # h (read by the path)
    title_size=512, feature_size=712, news_feature_size=800, bert_embed_size=512,
    query_vector_dim_large=400, news_encoder_size=600, long_short_term_method='ini',
    user_heads_num=8, num_heads_2=4, list_num_heads=8, kernel_sizes=3, kernel_sizes_2=[2, 4],
    num_filters=400, filter_nums_2=50, title_heads_num=6,
)

# additions of this implementation ----------------------------------------------------------------
_B200_KNOBS = dict(
    gemm_mode=1,       # 1: tcgen05 split-bf16 (bf16x3, fp32-grade); 2: tcgen05 plain bf16; 0: fp32 CUDA-core GEMMs
    dropout_seed=0,    # base Philox key; training step t uses dropout_seed + t
)


class Config(object):
    def __init__(self, model_name='NRMS', dataset='../MIND'):
        self.model_name = model_name
        for split in ('train', 'dev', 'test', 'small_train', 'small_dev'):
            setattr(self, split + '_path', '%s/%s/' % (dataset, split))
        self.log_path = './logs/' + model_name
        self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        for table in (_PATH_KNOBS, _INERT_KNOBS, _B200_KNOBS):
            for k, v in table.items():
                setattr(self, k, list(v) if isinstance(v, list) else v)

    def __nrms__(self):
        """Adds the model dimensions; the reference requires this call before the model is
        built (run_v0.py:50) and so do we (`Model.__init__` reads query_vector_dim)."""
        for k, v in _NRMS_KNOBS.items():
            setattr(self, k, list(v) if isinstance(v, list) else v)
        return self
