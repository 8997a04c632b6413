/* AudioStream.h — TEST INFRASTRUCTURE (oracle/_ref build).  The slice of the Teensy Audio
 * runtime the in-tree nodes touch: ref-counted blocks, receiveReadOnly, release. */
#ifndef ORACLE_SHIM_AUDIOSTREAM_H
#define ORACLE_SHIM_AUDIOSTREAM_H
#include "Arduino.h"

typedef struct audio_block_struct {
    uint8_t  ref_count;
    uint8_t  reserved1;
    uint16_t memory_pool_index;
    int16_t  data[AUDIO_BLOCK_SAMPLES];
} audio_block_t;

class AudioStream {
public:
    AudioStream(unsigned char ninput, audio_block_t **iqueue) : num_inputs(ninput), inputQueue(iqueue) {
        for (int i = 0; i < ninput; i++) iqueue[i] = NULL;
    }
    virtual ~AudioStream() {}
    virtual void update(void) = 0;
    /* harness side: queue a block on an input port (what AudioConnection + transmit() do) */
    void shim_feed(unsigned int index, const int16_t *samples) {
        audio_block_t *b = new audio_block_t;
        b->ref_count = 1; b->reserved1 = 0; b->memory_pool_index = 0;
        memcpy(b->data, samples, sizeof(b->data));
        if (inputQueue[index]) release(inputQueue[index]);
        inputQueue[index] = b;
    }
protected:
    audio_block_t *receiveReadOnly(unsigned int index = 0) {
        if (index >= num_inputs) return NULL;
        audio_block_t *in = inputQueue[index];
        inputQueue[index] = NULL;
        return in;
    }
    static void release(audio_block_t *block) {
        if (!block) return;
        if (block->ref_count > 1) block->ref_count--;
        else delete block;
    }
private:
    unsigned char num_inputs;
    audio_block_t **inputQueue;
};
#endif
