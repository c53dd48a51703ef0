"""Hyper-parameter surface of the reference (settings.Config, settings.py:27-211) without click/TensorFlow.

Field names and defaults are the reference's (`main.py train` options, main.py:12-54; cook-only fields from
main.py:100-147) so a config dict written for the reference constructs the same model here.
"""
import os

SLOTS = [
    'input_training_data_path', 'input_validation_data_path', 'input_previous_model_path', 'output_model_path', 'log_dir',
    'task', 'evaluate_sessions', 'enable_baseline', 'target', 'epochs', 'batch_size', 'learning_rate',
    'learning_rate_decay', 'training_step', 'validation_step', 'testing_impression', 'validation_impression', 'arch',
    'user_embedding_dim', 'title_filter_shape', 'title_shape', 'body_shape', 'window_size', 'hidden_dim', 'days', 'gain',
    'round', 'dropout', 'negative_samples', 'nonlocal_negative_samples', 'textual_embedding_dim',
    'textual_embedding_trainable', 'enable_pretrain_encoder', 'pretrain_encoder_trainable', 'name', 'pretrain_name',
    'debug', 'background', 'personal_embedding_dim', 'news_encoder', 'use_vertical', 'pipeline_input', 'body_sent_cnt',
    'body_sent_len', 'body_filter_shape', 'max_impression', 'max_impression_pos', 'max_impression_neg', 'score_model',
    'test_window_size', 'vertical_embedding_dim', 'subvertical_embedding_dim', 'lrd_on_epochs', 'id_keep',
    'use_generator', 'user_feature_size', 'doc_feature_size', 'use_vertical_type',
    # extensions of this implementation (not in the reference): engine precision and Keras recurrent activation
    'precision', 'recurrent_activation', 'sparse_user_adam',
]

TRAIN_DEFAULTS = dict(           # main.py:12-54
    task='UserEmbedding', arch='avg', round=6, days=30, epochs=10, batch_size=100, training_step=10000,
    validation_step=1000, validation_impression=1000, testing_impression=1000, learning_rate=0.001,
    learning_rate_decay=0.2, gain=1.0, window_size=10, dropout=0.2, negative_samples=4, hidden_dim=400,
    nonlocal_negative_samples=0, enable_baseline=False, title_filter_shape=(400, 3), title_shape=20, body_shape=200,
    user_embedding_dim=200, textual_embedding_dim=300, textual_embedding_trainable=False, debug=False, background=False,
    name='', pretrain_name='', enable_pretrain_encoder=False, pretrain_encoder_trainable=False,
    personal_embedding_dim=20, news_encoder='cnnatt', score_model='dot', body_sent_cnt=50, body_sent_len=30,
    body_filter_shape=(400, 3), max_impression=200, max_impression_pos=7, max_impression_neg=200, test_window_size=100,
    vertical_embedding_dim=10, subvertical_embedding_dim=20,
    input_training_data_path='.', input_validation_data_path='.', input_previous_model_path='.', output_model_path='.',
    log_dir='.', precision='auto', recurrent_activation='hard_sigmoid', sparse_user_adam=False,
    # cook-only options (main.py:100-147)
    id_keep=1.0, lrd_on_epochs=[1, 3], use_vertical=False, use_vertical_type='vs', use_generator=False,
)


class Config:
    __slots__ = SLOTS

    def __init__(self, config):
        merged = dict(TRAIN_DEFAULTS)
        merged.update(config)
        for k, v in merged.items():
            if not k.startswith('node'):            # settings.py:97-99
                setattr(self, k, v)

    # derived paths, settings.py:107-211
    @property
    def training_data_input(self):
        return os.path.join(self.input_training_data_path, 'ClickData.tsv')

    @property
    def testing_data_input(self):
        return os.path.join(self.input_training_data_path, 'TestData.tsv')

    @property
    def title_embedding_input(self):
        return os.path.join(self.input_training_data_path, 'Vocab.tsv')

    @property
    def doc_meta_input(self):
        return os.path.join(self.input_training_data_path, 'DocMeta.tsv')

    @property
    def model_output(self):
        return (os.path.join(self.output_model_path, 'model{}.json'.format(self.name)),
                os.path.join(self.output_model_path, 'model{}.pkl'.format(self.name)))

    @property
    def model_input(self):                         # settings.py:127-131
        return (os.path.join(self.input_previous_model_path, 'model{}.json'.format(self.pretrain_name)),
                os.path.join(self.input_previous_model_path, 'model{}.pkl'.format(self.pretrain_name)))

    @property
    def encoder_input(self):                       # settings.py:139-143 (--enable-pretrain-encoder)
        return (os.path.join(self.input_previous_model_path, 'encoder{}.json'.format(self.pretrain_name)),
                os.path.join(self.input_previous_model_path, 'encoder{}.pkl'.format(self.pretrain_name)))

    @property
    def encoder_output(self):                      # settings.py:145-149
        return (os.path.join(self.output_model_path, 'encoder{}.json'.format(self.name)),
                os.path.join(self.output_model_path, 'encoder{}.pkl'.format(self.name)))

    @property
    def log_output(self):
        return os.path.join(self.log_dir, 'log{}.txt'.format(self.name))

    @property
    def pipeline_inputs(self):
        """settings.py:169-174"""
        return (os.path.join(self.pipeline_input, 'docs.tsv'), os.path.join(self.pipeline_input, 'UserClick.tsv'),
                os.path.join(self.pipeline_input, 'userDocPair.tsv'))

    @property
    def pipeline_output(self):
        """settings.py:176-178"""
        return os.path.join(self.pipeline_input, 'score_' + self.name + '.tsv')

    @property
    def train_npz_input(self):
        return os.path.join(self.input_training_data_path, 'train_{}days_{}window.npz').format(self.days, self.window_size)

    @property
    def test_npz_input(self):
        return os.path.join(self.input_training_data_path, 'test_{}days_{}window.npz').format(self.days, self.window_size)
