// piper.hpp-compatible host surface over the piper_b200 engine.
//
// Same namespace, type names, fields and function signatures as the reference's public header
// (/root/reference/src/cpp/piper.hpp:20-128) so that its callers (src/cpp/main.cpp, src/cpp/test.cpp)
// compile against this file unchanged, with two substitutions:
//   * ModelSession holds a `pb200_voice*` (the C ABI of include/piper_b200.h) where the reference
//     holds Ort::Session / Ort::Env / Ort::SessionOptions (piper.hpp:78-85);
//   * text -> phonemes stays on the host behind a hook (PiperConfig::phonemizer) because espeak-ng /
//     piper-phonemize are external to the reference tree (CMakeLists.txt:63-71) and absent here; a
//     host that links piper-phonemize installs phonemize_eSpeak there.  Pre-phonemized input
//     (etc/test_sentences/*.jsonl: "phonemes" / "phoneme_ids") enters through phonemesToAudio /
//     phonemeIdsToAudio, which share all code with textToAudio below the phonemizer call.
#ifndef PIPER_B200_SHIM_PIPER_H_
#define PIPER_B200_SHIM_PIPER_H_

#include <cstdint>
#include <functional>
#include <map>
#include <optional>
#include <ostream>
#include <string>
#include <vector>

#include "mini_json.h"

struct pb200_voice;

namespace piper {

typedef int64_t SpeakerId;
typedef char32_t Phoneme;     // piper-phonemize: UTF-32 codepoint
typedef int64_t PhonemeId;    // backs an int64 tensor (piper.cpp:351-355)
using json = minijson::Value;

struct eSpeakConfig {
  std::string voice = "en-us";
};

// text (UTF-8), espeak voice -> one phoneme vector per sentence
using Phonemizer =
    std::function<void(const std::string&, const eSpeakConfig&, std::vector<std::vector<Phoneme>>&)>;

struct PiperConfig {
  std::string eSpeakDataPath;
  bool useESpeak = true;
  bool useTashkeel = false;
  std::optional<std::string> tashkeelModelPath;
  Phonemizer phonemizer;  // host hook (espeak-ng); unset => textToAudio throws, phonemesToAudio still works
};

enum PhonemeType { eSpeakPhonemes, TextPhonemes };

struct PhonemizeConfig {
  PhonemeType phonemeType = eSpeakPhonemes;
  std::optional<std::map<Phoneme, std::vector<Phoneme>>> phonemeMap;
  std::map<Phoneme, std::vector<PhonemeId>> phonemeIdMap;
  PhonemeId idPad = 0;
  PhonemeId idBos = 1;
  PhonemeId idEos = 2;
  bool interspersePad = true;
  eSpeakConfig eSpeak;
};

struct SynthesisConfig {
  float noiseScale = 0.667f;
  float lengthScale = 1.0f;
  float noiseW = 0.8f;
  int sampleRate = 22050;
  int sampleWidth = 2;
  int channels = 1;
  std::optional<SpeakerId> speakerId;
  float sentenceSilenceSeconds = 0.2f;
  std::optional<std::map<Phoneme, float>> phonemeSilenceSeconds;
};

struct ModelConfig {
  int numSpeakers = 1;
  std::optional<std::map<std::string, SpeakerId>> speakerIdMap;
};

struct ModelSession {
  pb200_voice* engine = nullptr;
  int device = 0;
  uint64_t noiseSeed = 0x9E3779B97F4A7C15ull;  // advanced per call: the reference's noise is unseeded
  ModelSession() = default;
  ModelSession(const ModelSession&) = delete;
  ModelSession& operator=(const ModelSession&) = delete;
  ~ModelSession();
};

struct SynthesisResult {
  double inferSeconds = 0;
  double audioSeconds = 0;
  double realTimeFactor = 0;
};

struct Voice {
  json configRoot;
  PhonemizeConfig phonemizeConfig;
  SynthesisConfig synthesisConfig;
  ModelConfig modelConfig;
  ModelSession session;
};

bool isSingleCodepoint(std::string s);
Phoneme getCodepoint(std::string s);
std::string getVersion();

void initialize(PiperConfig& config);
void terminate(PiperConfig& config);

// `useCuda` selects the CUDA device 0 engine; false is refused (this library has no CPU path).
void loadVoice(PiperConfig& config, std::string modelPath, std::string modelConfigPath, Voice& voice,
               std::optional<SpeakerId>& speakerId, bool useCuda);

// Phoneme ids -> int16 audio appended to audioBuffer (the reference's internal `synthesize`, piper.cpp:337-441)
void synthesize(std::vector<PhonemeId>& phonemeIds, SynthesisConfig& synthesisConfig, ModelSession& session,
                std::vector<int16_t>& audioBuffer, SynthesisResult& result);

// piper-phonemize's phonemes_to_ids: BOS, PAD, (ids(p), PAD)*, EOS; unknown phonemes are counted, not fatal
void phonemes_to_ids(const std::vector<Phoneme>& phonemes, const PhonemizeConfig& config,
                     std::vector<PhonemeId>& phonemeIds, std::map<Phoneme, std::size_t>& missingPhonemes);

void textToAudio(PiperConfig& config, Voice& voice, std::string text, std::vector<int16_t>& audioBuffer,
                 SynthesisResult& result, const std::function<void()>& audioCallback);
void textToWavFile(PiperConfig& config, Voice& voice, std::string text, std::ostream& audioFile,
                   SynthesisResult& result);

// Front doors for pre-phonemized input (same sentence / phrase / silence logic as textToAudio).
void phonemesToAudio(PiperConfig& config, Voice& voice, std::vector<std::vector<Phoneme>>& sentences,
                     std::vector<int16_t>& audioBuffer, SynthesisResult& result,
                     const std::function<void()>& audioCallback);
void phonemeIdsToWavFile(Voice& voice, std::vector<PhonemeId>& phonemeIds, std::ostream& audioFile,
                         SynthesisResult& result);
void writeWavHeader(int sampleRate, int sampleWidth, int channels, uint32_t numSamples, std::ostream& audioFile);

}  // namespace piper

#endif  // PIPER_B200_SHIM_PIPER_H_
