/* ref_tables.c — TEST INFRASTRUCTURE (oracle/_ref build).  Storage for the Teensy window tables
 * that analyze_fft256iq.h:33-50 declares extern; filled from the formula-generated tables
 * (pinned against the firmware image in tests/test_oracle_tables.py) before main(). */
#include <stdint.h>
#include <string.h>
#include "teensy_shim.h"

int16_t AudioWindowHanning256[256];
int16_t AudioWindowBlackmanNuttall256[256];   /* class default (analyze_fft256iq.h:56); the sketch overrides it, RadioDSP_SDR_RX.ino:144 */

__attribute__((constructor)) static void fill_windows(void)
{
    memcpy(AudioWindowHanning256, oracle_hanning256(), sizeof(AudioWindowHanning256));
}
