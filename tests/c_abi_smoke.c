/* c_abi_smoke.c — include/wam.h seen by a plain C compiler: the header is valid C, and the layouts a foreign-function
 * binding depends on (sizes and field offsets of every struct that crosses the boundary) are printed as JSON for
 * tests/test_lib_and_host.py to compare with the ctypes mirror in webaudio-modem_b200/_lib.py.  No calls, no GPU. */
#include <stddef.h>
#include <stdio.h>

#include "../include/wam.h"

#define S(T) printf("%s\"%s\": {\"size\": %zu", first++ ? ",\n " : " ", #T, sizeof(T))
#define F(T, f) printf(", \"%s\": %zu", #f, offsetof(T, f))
#define E() printf("}")

int main(void) {
  int first = 0;
  printf("{");
  S(wam_fsk_config); F(wam_fsk_config, sampleRate); F(wam_fsk_config, baudRate); F(wam_fsk_config, markFrequency);
  F(wam_fsk_config, spaceFrequency); F(wam_fsk_config, preamblePattern); F(wam_fsk_config, preambleLength);
  F(wam_fsk_config, sfdPattern); F(wam_fsk_config, sfdLength); F(wam_fsk_config, startBits); F(wam_fsk_config, stopBits);
  F(wam_fsk_config, parity); F(wam_fsk_config, syncThreshold); F(wam_fsk_config, agcEnabled);
  F(wam_fsk_config, preFilterBandwidth); F(wam_fsk_config, adaptiveThreshold); E();
  S(wam_fsk_status); F(wam_fsk_status, ready); F(wam_fsk_status, frameStarted); F(wam_fsk_status, globalSampleCounter);
  F(wam_fsk_status, receivedBitsLength); F(wam_fsk_status, byteBufferLength); F(wam_fsk_status, demodulationCalls);
  F(wam_fsk_status, syncDetections); F(wam_fsk_status, silenceThreshold); F(wam_fsk_status, totalSamplesProcessed);
  F(wam_fsk_status, eodEvents); F(wam_fsk_status, errorEvents); F(wam_fsk_status, configuredEvents); E();
  S(wam_pkt_result); F(wam_pkt_result, status); F(wam_pkt_result, sequence); F(wam_pkt_result, length);
  F(wam_pkt_result, payloadOffset); F(wam_pkt_result, crcReceived); F(wam_pkt_result, crcComputed);
  F(wam_pkt_result, bytesConsumed); E();
  S(wam_xmodem_rx_state); F(wam_xmodem_rx_state, expectedSequence); F(wam_xmodem_rx_state, retries);
  F(wam_xmodem_rx_state, done); F(wam_xmodem_rx_state, dataLen); F(wam_xmodem_rx_state, packetsReceived);
  F(wam_xmodem_rx_state, packetsDropped); E();
  S(wam_chunk_result); F(wam_chunk_result, samples); F(wam_chunk_result, isComplete); F(wam_chunk_result, samplesConsumed);
  F(wam_chunk_result, totalSamples); E();
  S(wam_fast_stats); E();
  printf("}\n");
  return 0;
}
