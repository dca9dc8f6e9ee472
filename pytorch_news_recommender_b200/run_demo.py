"""The reference's `run_demo.py` (lines 20-61) on this package: same literal file names under
`config.data_path`, same seeds, same objects — `Config('NRMS_V0_DEMO')`, `load_dataset`,
`MyDataset` + `DataLoader` (or `--device-batcher`: batches assembled on the GPU), `NRMS_V0`,
`train_demo`.

    python -m pytorch_news_recommender_b200.run_demo --data-path ./data_processed/ [--synthetic]

`--synthetic` first writes MIND-shaped synthetic files with those names (the real MIND dumps are
not shipped with the reference and there is no network)."""
from __future__ import annotations

import argparse
import os

import torch
from torch.utils.data import DataLoader

from . import synthetic
from .config import Config
from .data_handler import DeviceBatcher, MyDataset, load_dataset
from .model import NRMS_V0
from .train_eval import train_demo


def main(argv=None):
    ap = argparse.ArgumentParser(description='MIND')
    ap.add_argument('--data-path', default='./data_processed/')
    ap.add_argument('--synthetic', action='store_true', help='write synthetic demo files into --data-path first')
    ap.add_argument('--device-batcher', action='store_true', help='assemble batches on the GPU instead of DataLoader workers')
    ap.add_argument('--batch-size', type=int, default=256)
    ap.add_argument('--epochs', type=int, default=None)
    ap.add_argument('--workers', type=int, default=6)
    args = ap.parse_args(argv)

    print('current: uid', os.getpid())
    torch.manual_seed(42)
    torch.cuda.manual_seed_all(42)
    model_name = 'NRMS_V0_DEMO'
    config = Config(model_name)
    config.__nrms__()
    config.batch_size = args.batch_size
    print(model_name, config.batch_size)
    config.mode = 'demo'
    config.word_embedding_pretrained = 'demo_word_embedding.npz'
    config.data_path = args.data_path if args.data_path.endswith('/') else args.data_path + '/'
    if args.epochs is not None:
        config.num_epochs = args.epochs
    if args.synthetic:
        synthetic.write_demo_files(config.data_path, config)

    train_list = load_dataset(config, 'small_train.pkl', config.data_path, _type=0)
    dev_list = load_dataset(config, 'small_dev.pkl', config.data_path, _type=1)
    if args.device_batcher:
        train_iter = DeviceBatcher(config, train_list, type=0, batch_size=config.batch_size, shuffle=True)
        dev_iter = DeviceBatcher(config, dev_list, type=1, batch_size=512, shuffle=False)
    else:
        train_iter = DataLoader(dataset=MyDataset(config, train_list, type=0), batch_size=config.batch_size,
                                num_workers=args.workers, drop_last=False, shuffle=True, pin_memory=False)
        dev_iter = DataLoader(dataset=MyDataset(config, dev_list, type=1), batch_size=512,
                              num_workers=args.workers, drop_last=False, shuffle=False, pin_memory=False)
    print("dev_data nums:::", len(dev_list))
    model = NRMS_V0(config).to(config.device)
    print(model.parameters)
    return train_demo(config, model, train_iter, dev_iter)


if __name__ == '__main__':
    main()
