"""Job launcher with the reference's surface (dev/py/run_local_ofdm.py, dev/py/locals.py).

The reference builds ``--flag=value`` command lines for ofdmreceiver_np.py / ofdmreceiver_np_mp.py
per (longcp, modulation, cp) job and shells out, skipping a job when its result CSV exists
(run_local_ofdm.py:74-90, 101-114).  This launcher builds the SAME flag strings and runs the jobs
in-process on the GPU(s): each ofdmreceiver_np job trains the basic receiver and sweeps its BER, each
ofdmreceiver_np_mp job transfer-learns the equalizer in front of that receiver and runs the cross-channel test; under
torchrun every job's SNR x channel grid is sharded over the ranks.
  python -m dl_ofdm_b200.run_local_ofdm --awgn=True
"""
from __future__ import annotations

import argparse
import os
import sys


def job_list(awgn=True, token='OFDM_Dense3', batchsize=512, nfft=64, ebno=5.0, learning=0.001, mobile=True):
    """-> (save_dir, result_dir, [(script, flag string, result csv)]) with the reference's strings and order
    (run_local_ofdm.py:30-118).  Reference quirks kept: save_dir / result_dir are computed ONCE before the loops (from
    longcp = 'True'), and the tokens do not carry longcp, so the second pass of the outer longcp loop finds the result
    CSVs of the first one and skips."""
    mobile_str = '_mobile' if mobile else ''
    save_dir = './ofdm_lte_ext_%s_longcp%s/' % (nfft, mobile_str)
    result_dir = './test_ext_%s_long_cross%s' % (nfft, mobile_str)
    jobs = []
    for longcp in ('False', 'True'):
        if awgn:
            for nbits in (1, 2, 3, 4):                                   # reversed([4,3,2,1])
                snr = float(ebno * nbits)
                for cp in ('False', 'True'):
                    flags = ('--channel=%s --save_dir=%s --early_stop=200 --nfilter=%d --batch_size=%d --max_epoch_num=%d '
                             '--cp=%s --nfft=%d --longcp=%s ' % ('AWGN', save_dir, nfft, batchsize, 1200 * nbits, cp, nfft,
                                                                 longcp))
                    tok = '%s_%dmod_snr%d_cp%s' % (token, nbits, int(snr), cp)
                    flags += '--SNR=%.2f --nbits=%d --token=%s' % (snr, nbits, tok)
                    jobs.append(('ofdmreceiver_np', flags, 'Test_DCCN_%s_AWGN.csv' % tok))
        nbits, opt = 1, 0
        snr = float(ebno * nbits)
        for cp in ('True', 'False'):
            flags = ('--channel=%s --save_dir=%s --init_learning=%.4f --early_stop=200 --nfilter=%d --batch_size=%d '
                     '--max_epoch_num=%d --cp=%s --nfft=%d --longcp=%s --opt=%d --mobile=%s '
                     % ('mixRayleigh', save_dir, learning, nfft, batchsize, 4000 * nbits, cp, nfft, longcp, opt, mobile))
            tok = '%s_%dmod_snr%d_cp%s' % (token, nbits, int(snr), cp)
            flags += '--SNR=%.2f --nbits=%d --token=%s' % (snr, nbits, tok)
            # the reference tests for ..._test_chan_Custom.csv (run_local_ofdm.py:107) although its mobile runs write
            # ..._Custom_mobile.csv (_mp.py:98-99), so with --mobile=True its skip rule never fires; both names are
            # accepted here
            jobs.append(('ofdmreceiver_np_mp', flags, 'Test_DCCN_%s_Equalizer%d_mixRayleigh_test_chan_Custom.csv' % (tok, opt)))
    return save_dir, result_dir, jobs


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--awgn', type=lambda v: str(v).lower() in ('1', 'true'), default=True)
    ap.add_argument('--dry_run', action='store_true', help='print the job list and exit')
    ap.add_argument('--max_epoch_num', type=int, default=0, help='override every job\'s --max_epoch_num (short runs)')
    args, _ = ap.parse_known_args(argv)
    from . import ofdmreceiver_np, ofdmreceiver_np_mp
    mods = {'ofdmreceiver_np': ofdmreceiver_np, 'ofdmreceiver_np_mp': ofdmreceiver_np_mp}
    save_dir, result_dir, jobs = job_list(args.awgn)
    if not args.dry_run:
        for folder in (save_dir, result_dir):
            os.makedirs(folder, exist_ok=True)
    from .sweep import init_distributed, barrier
    rank, _ = init_distributed()
    for script, flags, csv in jobs:
        alts = [csv, csv[:-4] + '_mobile.csv']
        if any(os.path.exists(c) or os.path.exists(os.path.join(result_dir, c)) for c in alts):   # resume rule (:82-86, :110-114)
            print('skip (result exists):', csv)
            continue
        if args.max_epoch_num:
            flags += ' --max_epoch_num=%d' % args.max_epoch_num
        print('python -u %s.py %s' % (script, flags))
        if not args.dry_run:
            try:
                mods[script].main(flags.split())       # trains (basic receiver / equalizer), then runs the BER test
            except FileNotFoundError as e:
                print('  ->', e)
                continue
            barrier()                                  # every rank is past the job before its files move
            pre = csv.split('_test_chan_')[0] if '_test_chan_' in csv else csv[:-4]
            if rank == 0:                              # rank 0 wrote the CSVs (test_model / test_model_cross)
                for f in os.listdir('.'):
                    if f.startswith(pre) and f.endswith('.csv'):
                        os.replace(f, os.path.join(result_dir, f))
            barrier()


if __name__ == '__main__':
    main(sys.argv[1:])
