// brb_policy_layout.h — the flat fp32 parameter block of the PPO actor-critic (SB3 state-dict order, ppo.py pack_params):
//   pi.W1[64][6] pi.b1[64] pi.W2[64][64] pi.b2[64] | vf.W1[64][6] vf.b1[64] vf.W2[64][64] vf.b2[64] |
//   action_net.W[2][64] action_net.b[2] | value_net.W[1][64] value_net.b[1] | log_std[2]
#ifndef BRB_POLICY_LAYOUT_H
#define BRB_POLICY_LAYOUT_H
#define PH 64                       // hidden width
#define PIN 6                       // observation size
#define TOWER (PH * PIN + PH + PH * PH + PH)
#define OFF_PI 0
#define OFF_VF TOWER
#define OFF_AW (2 * TOWER)
#define OFF_AB (OFF_AW + 2 * PH)
#define OFF_VW (OFF_AB + 2)
#define OFF_VB (OFF_VW + PH)
#define OFF_LS (OFF_VB + 1)
#define NPARAM (OFF_LS + 2)         // 9,413 = BRB_POLICY_NPARAM
#endif
