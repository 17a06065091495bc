"""End-to-end evidence for the training path: train the basic receiver (BPSK, AWGN, the launcher's first job) from
glorot-uniform variables on the GPU with the reference's schedule, then sweep BER over SNR.  Prints the curve next to
the v1 checkpoint's curve from BASELINE.md section 2 (the reference's own trained BPSK receiver; other frame layout --
8 symbols, scattered pilots -- so a ballpark comparison, not a parity target)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dl_ofdm_b200.flags import Flags
from dl_ofdm_b200.ofdm import ofdm_tx
from dl_ofdm_b200.ofdmreceiver_np import train_receiver, test_model
EPOCHS = int(os.environ.get('EPOCHS', 150))
V1_1B_CPT = {-10: 2.951e-01, -2: 8.944e-02, 0: 4.606e-02, 2: 1.744e-02, 3: 9.039e-03, 5: 1.670e-03, 7: 1.658e-04, 9: 1.359e-06}   # BASELINE.md section 2, column '1b cpT'
out = os.environ.get('OUT', '/tmp/train_rx_curve') + '/'
os.makedirs(out, exist_ok=True)
FLAGS = Flags(nbits=1, channel='AWGN', SNR=5.0, batch_size=512, msg_length=100800, token='OFDM_Dense3_1mod_snr5_cpTrue',
              save_dir=out, precision='parity', early_stop=200, cp=True, longcp=True)
ofdmobj = ofdm_tx(FLAGS)
t0 = time.time()
session, hist = train_receiver(FLAGS, ofdmobj, max_epoch_num=EPOCHS, log=lambda *a: None)
dt = time.time() - t0
print('trained %d epochs (%d Adam steps) in %.1f s; train loss %.4f -> %.4f, test BER @5 dB %.5f -> %.5f, minibatch %d -> %d frames'
      % (len(hist), hist[-1]['global_step'], dt, hist[0]['train_loss'], hist[-1]['train_loss'], hist[0]['test_ber'],
         hist[-1]['test_ber'], 512 // 7, hist[-1]['batch']))
rows = test_model(FLAGS, out + FLAGS.token, ofdmobj, session=session, frame_cnt=20000, snrs=range(-10, 11), out_dir=out)
print('SNR dB   BER (trained here, dev frame)   v1 checkpoint (BASELINE.md, v1 frame)')
for r in rows:
    s = int(r['SNR'])
    print('%4d     %.4e                      %s' % (s, r['BER'], ('%.3e' % V1_1B_CPT[s]) if s in V1_1B_CPT else ''))
session.close()
