/* Audio.h — TEST INFRASTRUCTURE (oracle/_ref build).  AudioRecordQueue / AudioPlayQueue as the
 * loop()-context code of RDSP_convolutional.h:231-245,342-350 uses them. */
#ifndef ORACLE_SHIM_AUDIO_H
#define ORACLE_SHIM_AUDIO_H
#include "AudioStream.h"
#include <deque>
#include <vector>

class AudioRecordQueue {
public:
    AudioRecordQueue() : enabled(false) {}
    void begin(void) { enabled = true; }
    int available(void) { return (int)q.size(); }
    int16_t *readBuffer(void) { return q.empty() ? NULL : q.front().data(); }
    void freeBuffer(void) { if (!q.empty()) q.pop_front(); }
    /* harness side: one update() tick delivers one block */
    void shim_push(const int16_t *s) { if (enabled) q.emplace_back(s, s + AUDIO_BLOCK_SAMPLES); }
private:
    bool enabled;
    std::deque<std::vector<int16_t> > q;
};

class AudioPlayQueue {
public:
    int16_t *getBuffer(void) { pending.assign(AUDIO_BLOCK_SAMPLES, 0); return pending.data(); }
    void playBuffer(void) { q.push_back(pending); }
    /* harness side */
    int shim_available(void) { return (int)q.size(); }
    void shim_pop(int16_t *dst) { memcpy(dst, q.front().data(), AUDIO_BLOCK_SAMPLES * sizeof(int16_t)); q.pop_front(); }
private:
    std::vector<int16_t> pending;
    std::deque<std::vector<int16_t> > q;
};
#endif
