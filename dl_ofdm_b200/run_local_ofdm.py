"""Job launcher with the reference's surface (dev/py/run_local_ofdm.py, dev/py/locals.py).

The reference builds ``--flag=value`` command lines for ofdmreceiver_np.py / ofdmreceiver_np_mp.py
per (longcp, modulation, cp) job and shells out, skipping a job when its result CSV exists
(run_local_ofdm.py:74-90, 101-114).  This launcher builds the SAME flag strings and runs the jobs
in-process on the GPU(s); under torchrun every job's SNR x channel grid is sharded over the ranks.
  python -m dl_ofdm_b200.run_local_ofdm --awgn=True
"""
from __future__ import annotations

import argparse
import os
import sys


def job_list(awgn=True, token='OFDM_Dense3', batchsize=512, nfft=64, ebno=5.0, save_dir='./output/'):
    """-> [(script, flag string, result csv)] in the reference's order (run_local_ofdm.py:30-118)."""
    jobs = []
    if awgn:
        for longcp in (False, True):
            for nbits in (4, 3, 2, 1):
                for cp in (False, True):
                    tok = '%s_%dmod_cp%s_longcp%s' % (token, nbits, cp, longcp)
                    flags = ('--save_dir=%s --token=%s --nbits=%d --batch_size=%d --nfft=%d --nfilter=%d --SNR=%.1f '
                             '--channel=AWGN --cp=%s --longcp=%s --early_stop=100 --test=True'
                             % (save_dir, tok, nbits, batchsize, nfft, nfft, ebno * nbits, cp, longcp))
                    jobs.append(('ofdmreceiver_np', flags, 'Test_DCCN_%s_AWGN.csv' % tok))
    for cp in (False, True):
        tok = '%s_1mod_cp%s_longcpTrue' % (token, cp)
        flags = ('--save_dir=%s --token=%s --nbits=1 --batch_size=%d --nfft=%d --nfilter=%d --channel=mixRayleigh '
                 '--cp=%s --longcp=True --opt=0 --mobile=False --test=True' % (save_dir, tok, batchsize, nfft, nfft, cp))
        jobs.append(('ofdmreceiver_np_mp', flags, 'Test_DCCN_%s_Equalizer0_mixRayleigh_test_chan_EPA.csv' % tok))
    return jobs


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument('--awgn', type=lambda v: str(v).lower() in ('1', 'true'), default=True)
    ap.add_argument('--dry_run', action='store_true', help='print the job list and exit')
    args, _ = ap.parse_known_args(argv)
    from . import ofdmreceiver_np, ofdmreceiver_np_mp
    mods = {'ofdmreceiver_np': ofdmreceiver_np, 'ofdmreceiver_np_mp': ofdmreceiver_np_mp}
    for script, flags, csv in job_list(args.awgn):
        if os.path.exists(csv):                       # resume rule of the reference (:82-86, :110-114)
            print('skip (result exists):', csv)
            continue
        print('python -u %s.py %s' % (script, flags))
        if not args.dry_run:
            try:
                mods[script].main(flags.split())
            except FileNotFoundError as e:
                print('  ->', e)


if __name__ == '__main__':
    main(sys.argv[1:])
